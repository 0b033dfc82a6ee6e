"""Bit-exact port of the scheduler's integer memory-packing arithmetic (oracle; see oracle/__init__.py).

These are the functions of SURVEY.md §8a rows a7/a8 that decide WHERE a ModelInstance may live; the new
runtime must honour their outputs (gpu_index, model_memory_requirement, --gpu-memory-utilization,
--max-num-seqs), so tests pin them against the reference's own known-answer tests.
"""


def vllm_memory_utilization_ratio(per_gpu_memory: int, model_memory_requirement: int) -> float:
    """api/pkg/scheduler/runner.go:1187-1219 calculateVLLMMemoryUtilizationRatio."""
    if per_gpu_memory == 0:
        return 0.8
    r = float(model_memory_requirement) / float(per_gpu_memory)
    if r < 0.01:
        r = 0.01
    elif r > 0.99:
        r = 0.99
    return r


def ratio_arg(r: float) -> str:
    """runner.go:1225 fmt.Sprintf("%.2f", ratio)."""
    return "%.2f" % r


def substitute_vllm_args(args, per_gpu_memory, model_memory_requirement):
    """runner.go:1222-1259 substituteVLLMArgsPlaceholders."""
    rs = ratio_arg(vllm_memory_utilization_ratio(per_gpu_memory, model_memory_requirement))
    out = list(args)
    has = False
    for i, a in enumerate(out):
        if a == "{{.DynamicMemoryUtilizationRatio}}":
            out[i] = rs
        elif a == "--gpu-memory-utilization":
            has = True
    if not has:
        out += ["--gpu-memory-utilization", rs]
    return out


def single_gpu_fit(total, allocated_per_gpu, need):
    """api/pkg/scheduler/global_allocator.go:349-452 (no-eviction branch): GPUs where total-allocated >= need,
    ranked by cost = usedGB after placement (global_allocator.go:667-680, runner load term equal within a runner)."""
    plans = []
    for idx, alloc in sorted(allocated_per_gpu.items()):
        free = total[idx] - alloc if total[idx] >= alloc else 0
        if free >= need:
            plans.append((int((alloc + need) // (1024 ** 3)), idx))
    plans.sort()
    return [i for _, i in plans]


def multi_gpu_split(need, n):
    """global_allocator.go:455-549: even split need/n per GPU."""
    return need // n


def pick_best_warm_slot(slots):
    """api/pkg/scheduler/scheduler.go:1958-2009: fewest active requests, then least-loaded runner, then most recent.
    slots: list of dicts {id, active, runner_load, last_activity}; returns the chosen id (deterministic part)."""
    best = min(slots, key=lambda s: (s["active"], s["runner_load"], -s["last_activity"]))
    return best["id"]


# ---- allocation planning with eviction (single runner view) ------------------------------------------------------
# Restatement of GlobalAllocator.PlanAllocation for one runner: api/pkg/scheduler/global_allocator.go:349-452 (single-GPU
# plans), :455-549 (multi-GPU plans), :551-587 (cheapest plan wins), :589-660 (evictable = stale slots of OTHER
# model/runtime/lora on that GPU, oldest first; evict the fewest that free enough), :667-707 (costs) and
# api/pkg/scheduler/runner.go:658-746 (allocated bytes per GPU: whole model on its single GPU, model/n on each of n GPUs).
# Everything is integer arithmetic on byte counts: the B200 runtime must fit what these plans promise, bit for bit.
GIB = 1024 ** 3


def allocated_per_gpu(slots):
    """runner.go:658-746.  slots: dicts {id, model, runtime, lora, gpus: [int, ...], memory, stale, last_activity}."""
    out = {}
    for s in slots:
        g = s["gpus"]
        if len(g) > 1:
            for i in g:
                out[i] = out.get(i, 0) + s["memory"] // len(g)
        elif len(g) == 1:
            out[g[0]] = out.get(g[0], 0) + s["memory"]
    return out


def _evictable(slots, gpu, work):
    same = lambda s: (s["model"], s["runtime"], s.get("lora", "")) == (work["model"], work["runtime"], work.get("lora", ""))
    ev = [s for s in slots if not same(s) and gpu in s["gpus"] and s["stale"]]
    return sorted(ev, key=lambda s: s["last_activity"])  # oldest first


def _select_for_eviction(evictable, needed):
    chosen, freed = [], 0
    for s in evictable:
        if freed >= needed:
            break
        chosen.append(s)
        freed += s["memory"]
    return chosen


def plan_allocation(gpus, slots, work):
    """gpus: [(index, total_bytes)] in runner order; work: {model, runtime, lora?, memory}.
    Returns the cheapest plan {gpus, memory_per_gpu, evict: [slot ids], cost, multi} or None ("no viable allocation
    plans")."""
    need = work["memory"]
    alloc = allocated_per_gpu(slots)
    runner_penalty = sum(s["memory"] for s in slots) // GIB
    plans = []
    for idx, total in gpus:                                     # single-GPU plans
        if need > total:
            continue
        free = total - alloc.get(idx, 0)
        if free >= need:
            plans.append({"gpus": [idx], "memory_per_gpu": need, "evict": [], "multi": False,
                          "cost": (total - free) // GIB + runner_penalty})
        else:
            ev = _evictable(slots, idx, work)
            if free + sum(s["memory"] for s in ev) >= need:
                chosen = _select_for_eviction(ev, need - free)
                plans.append({"gpus": [idx], "memory_per_gpu": need, "evict": [s["id"] for s in chosen], "multi": False,
                              "cost": 100 * len(chosen) + (total - free) // GIB + runner_penalty})
    for n in range(2, len(gpus) + 1):                           # multi-GPU plans: even split over the first n that fit
        per = need // n
        viable, evict, ok = [], [], True
        for idx, total in gpus:
            free = total - alloc.get(idx, 0)
            if free >= per:
                viable.append(idx)
            else:
                ev = _evictable(slots, idx, work)
                if free + sum(s["memory"] for s in ev) >= per:
                    evict += _select_for_eviction(ev, per - free)
                    viable.append(idx)
                else:
                    ok = False
                    break
            if len(viable) >= n:
                break
        if ok and len(viable) >= n:
            plans.append({"gpus": viable[:n], "memory_per_gpu": per, "evict": [s["id"] for s in evict], "multi": True,
                          "cost": 100 * len(evict) + 1000 * n + runner_penalty})
    if not plans:
        return None
    return min(enumerate(plans), key=lambda ip: (ip[1]["cost"], ip[0]))[1]   # stable: first of the cheapest
