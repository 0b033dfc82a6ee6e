"""Known-answer vectors lifted from the reference's own tests (the only golden values it holds for this path):
api/pkg/scheduler/runner_test.go:88-186 (TestCalculateVLLMMemoryUtilizationRatio) and :188-260 (args substitution)."""
import pytest

from oracle import scheduler_ref as S

GB = 1024 ** 3


@pytest.mark.parametrize("gpu,model,want,delta", [
    (80 * GB, 8 * GB, 0.10, 0.01), (80 * GB, 1 * GB, 0.0125, 0.001), (24 * GB, 16 * GB, 0.67, 0.01),
    (24 * GB, 20 * GB, 0.833, 0.01), (24 * GB, 8 * GB, 0.33, 0.01),
])
def test_ratio_in_delta(gpu, model, want, delta):
    assert abs(S.vllm_memory_utilization_ratio(gpu, model) - want) <= delta


def test_ratio_exact_cases():
    assert S.vllm_memory_utilization_ratio(0, 8 * GB) == 0.8          # no GPU info -> fallback
    assert S.vllm_memory_utilization_ratio(8 * GB, 10 * GB) == 0.99   # model > GPU -> clamp
    assert S.vllm_memory_utilization_ratio(24 * GB, 24 * GB) == 0.99
    assert S.vllm_memory_utilization_ratio(80 * GB, 1) == 0.01


def test_args_substitution():
    a = S.substitute_vllm_args(["--gpu-memory-utilization", "{{.DynamicMemoryUtilizationRatio}}", "--max-model-len", "8192"],
                               80 * GB, 8 * GB)
    assert a == ["--gpu-memory-utilization", "0.10", "--max-model-len", "8192"]
    b = S.substitute_vllm_args(["--max-model-len", "8192"], 80 * GB, 40 * GB)
    assert b[-2:] == ["--gpu-memory-utilization", "0.50"]
    assert S.substitute_vllm_args([], 0, 1) == ["--gpu-memory-utilization", "0.80"]


def test_single_gpu_fit_and_routing():
    total = {0: 80 * GB, 1: 80 * GB}
    assert S.single_gpu_fit(total, {0: 70 * GB, 1: 10 * GB}, 20 * GB) == [1]
    assert S.single_gpu_fit(total, {0: 10 * GB, 1: 30 * GB}, 20 * GB) == [0, 1]
    assert S.single_gpu_fit(total, {0: 70 * GB, 1: 70 * GB}, 20 * GB) == []
    assert S.multi_gpu_split(40 * GB, 2) == 20 * GB
    slots = [{"id": "a", "active": 2, "runner_load": 0, "last_activity": 5},
             {"id": "b", "active": 1, "runner_load": 9, "last_activity": 1},
             {"id": "c", "active": 1, "runner_load": 3, "last_activity": 0}]
    assert S.pick_best_warm_slot(slots) == "c"


def _slot(i, model, gpus, mem_gb, stale=False, age=0):
    return {"id": i, "model": model, "runtime": "vllm", "gpus": gpus, "memory": mem_gb * GB, "stale": stale,
            "last_activity": -age}


def test_overscheduling_prevention_scenario():
    """api/pkg/scheduler/global_allocator_test.go:452-521: two 70 GB models must land on different 80 GB GPUs; a 200 GB
    model fits nowhere, not even split ("no viable allocation plans")."""
    gpus = [(0, 80 * GB), (1, 80 * GB)]
    w70 = {"model": "medium-vllm:20b", "runtime": "vllm", "memory": 70 * GB}
    p1 = S.plan_allocation(gpus, [], w70)
    assert p1["gpus"] == [0] and not p1["evict"] and not p1["multi"]
    slots = [_slot("s1", "medium-vllm:20b", p1["gpus"], 70)]
    p2 = S.plan_allocation(gpus, slots, w70)
    assert p2["gpus"] == [1] and not p2["evict"]
    slots.append(_slot("s2", "medium-vllm:20b", p2["gpus"], 70))
    assert S.plan_allocation(gpus, slots, {"model": "huge-vllm:200b", "runtime": "vllm", "memory": 200 * GB}) is None
    assert S.allocated_per_gpu(slots) == {0: 70 * GB, 1: 70 * GB}
    assert all(v <= 80 * GB for v in S.allocated_per_gpu(slots).values())


def test_eviction_scenario():
    """global_allocator_test.go:342-450: both GPUs hold a stale 70 GB slot; a 45 GB model needs exactly one eviction
    (the oldest stale slot of the GPU the cheapest plan picks); slots of the SAME model are never evicted for it."""
    gpus = [(0, 80 * GB), (1, 80 * GB)]
    slots = [_slot("stale-1", "stale-model-1", [0], 70, stale=True, age=3600),
             _slot("stale-2", "stale-model-2", [1], 70, stale=True, age=3600)]
    plan = S.plan_allocation(gpus, slots, {"model": "medium-vllm:20b", "runtime": "vllm", "memory": 45 * GB})
    assert plan is not None and len(plan["evict"]) == 1 and plan["evict"][0] in ("stale-1", "stale-2")
    assert plan["gpus"] == [0] and plan["cost"] == 100 + 70 + 140          # eviction + used GB on the GPU + runner load GB
    fresh = [_slot("a", "m1", [0], 70), _slot("b", "m2", [1], 70)]         # running, not stale: nothing may be evicted
    assert S.plan_allocation(gpus, fresh, {"model": "x", "runtime": "vllm", "memory": 45 * GB}) is None
    same = [_slot("a", "x", [0], 70, stale=True), _slot("b", "x", [1], 70, stale=True)]
    assert S.plan_allocation(gpus, same, {"model": "x", "runtime": "vllm", "memory": 45 * GB}) is None


def test_single_gpu_preferred_over_split_and_split_arithmetic():
    """global_allocator.go:683-693: a multi-GPU plan carries +1000 per GPU, so it only wins when no single GPU can take
    the model; the split is the integer quotient (runner.go:697-702 accounts it the same way)."""
    gpus = [(i, 80 * GB) for i in range(4)]
    p = S.plan_allocation(gpus, [], {"model": "m", "runtime": "vllm", "memory": 60 * GB})
    assert p["gpus"] == [0] and not p["multi"]
    p = S.plan_allocation(gpus, [], {"model": "big", "runtime": "vllm", "memory": 150 * GB + 1})
    assert p["multi"] and p["gpus"] == [0, 1] and p["memory_per_gpu"] == (150 * GB + 1) // 2 and p["cost"] == 2000
    slots = [_slot("s", "big", p["gpus"], 150)]
    assert S.allocated_per_gpu(slots) == {0: 75 * GB, 1: 75 * GB}
    # least-allocated GPU wins among single-GPU plans (cost = used GB after the runner penalty, equal for all)
    slots = [_slot("a", "m1", [0], 30), _slot("b", "m2", [1], 10), _slot("c", "m3", [2], 20)]
    assert S.plan_allocation(gpus, slots, {"model": "n", "runtime": "vllm", "memory": 20 * GB})["gpus"] == [3]
    assert S.plan_allocation(gpus[:3], slots, {"model": "n", "runtime": "vllm", "memory": 20 * GB})["gpus"] == [1]
