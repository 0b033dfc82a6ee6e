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
