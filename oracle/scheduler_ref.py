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
