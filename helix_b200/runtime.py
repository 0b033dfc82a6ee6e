"""Host-side mirror of the reference's `runner.Runtime` interface (api/pkg/runner/slot.go:46-57) for the
B200 engine.  Same method set and meaning as the Go interface — Start/Stop/PullModel/Warm/ListModels/
Version/Status/Runtime/URL/CommandLine — so the Go shim in integration/go/ is a line-for-line cgo
rendering of this file.  Plug-in "Option A" of SURVEY.md §8b: the slot keeps `runtime:"vllm"` on the wire
and the scheduler's vLLM-style args arrive unchanged (`--gpu-memory-utilization`, `--max-num-seqs`,
`--max-model-len`, `--task embed`; api/pkg/scheduler/runner.go:1187-1259,1344-1397).
"""
import time
from dataclasses import dataclass, field
from typing import Callable, List, Optional

from . import configs
from .engine import Engine, EngineConfig, HBError, ModelDesc, Sampling, memory_estimate

VERSION = "helix-b200/0.2 (abi 2)"
DEFAULT_MAX_NUM_SEQS = 256  # types/memory.go:11 (vLLM default concurrency)


@dataclass
class ParsedArgs:
    gpu_memory_utilization: Optional[float] = None
    max_num_seqs: int = DEFAULT_MAX_NUM_SEQS
    max_model_len: Optional[int] = None
    task_embed: bool = False
    enable_prefix_caching: bool = True       # vLLM V1's default; multi-turn Helix sessions resend the whole conversation
    max_num_batched_tokens: Optional[int] = None
    unknown: List[str] = field(default_factory=list)


def parse_vllm_args(args: List[str]) -> ParsedArgs:
    """The subset of vLLM CLI flags the scheduler emits for a slot (vllm_runtime.go:705-762)."""
    out = ParsedArgs()
    i = 0
    while i < len(args):
        a = args[i]
        nxt = args[i + 1] if i + 1 < len(args) else None
        if a == "--gpu-memory-utilization" and nxt is not None:
            out.gpu_memory_utilization = float(nxt); i += 2
        elif a == "--max-num-seqs" and nxt is not None:
            out.max_num_seqs = int(nxt); i += 2
        elif a == "--max-model-len" and nxt is not None:
            out.max_model_len = int(nxt); i += 2
        elif a == "--task" and nxt is not None:
            out.task_embed = (nxt == "embed"); i += 2
        elif a == "--max-num-batched-tokens" and nxt is not None:
            out.max_num_batched_tokens = int(nxt); i += 2
        elif a in ("--enable-prefix-caching", "--no-enable-prefix-caching"):
            out.enable_prefix_caching = not a.startswith("--no-"); i += 1
        elif a.startswith("--") and nxt is not None and not nxt.startswith("--"):
            out.unknown += [a, nxt]; i += 2
        else:
            out.unknown.append(a); i += 1
    return out


def memory_budget(model_memory_requirement: int, per_gpu_memory: int, gpu_memory_utilization: Optional[float]) -> int:
    """Bytes the engine may use. The scheduler packs slots by `model_memory_requirement`
    (scheduler/model_allocation.go:39-60); the ratio flag is the same number divided by per-GPU memory,
    clamped to [0.01, 0.99] and printed with two decimals (runner.go:1187-1225) — prefer the exact bytes."""
    if model_memory_requirement:
        return int(model_memory_requirement)
    if gpu_memory_utilization and per_gpu_memory:
        return int(per_gpu_memory * gpu_memory_utilization)
    return 0


MODEL_CATALOGUE = {
    # slot.Model -> (description, is_embedding). Random-init shapes: no checkpoints exist offline.
    "meta-llama/Meta-Llama-3-8B-Instruct": (configs.llama3_8b, False),
    "meta-llama/Llama-3.2-1B-Instruct": (configs.llama32_1b, False),
    "BAAI/bge-base-en-v1.5": (configs.bge_base, True),
    # the reference's default embedding model (api/pkg/model/models.go:421-433): text backbone, last-token pooling
    "MrLight/dse-qwen2-2b-mrl-v1": (configs.dse_qwen2_2b, True),
}


def memory_estimation(request: dict, runner_id: str = "", catalogue=None) -> dict:
    """The runner's POST /memory-estimate for models served by this runtime (scope row F3).

    The reference answers it by parsing a GGUF and calling Ollama's layer estimator for 1/2/4/8 synthetic 80 GB GPUs
    (api/pkg/runner/memory_estimation_handlers.go:36-327; request/response: api/pkg/types/memory.go:16-50).  Here it is a
    closed form — `hb_memory_estimate`: weight arena + max_seqs full contexts of paged KV + step workspace — and there is
    one configuration, `single_gpu`: the runtime places a ModelInstance on one GPU (replicas, no tensor split), which is
    also what the allocator prefers (multi-GPU plans cost +1000 per GPU, global_allocator.go:683-693).  `num_parallel`
    is the slot's concurrency (--max-num-seqs); like the reference, KV is sized for num_parallel x context_length."""
    t0 = time.time()
    cat = catalogue or MODEL_CATALOGUE
    name = request.get("model_name", "")
    resp = {"success": False, "model_name": name, "model_path": "", "architecture": "", "block_count": 0,
            "configurations": [], "response_time_ms": 0, "runner_id": runner_id}
    if name not in cat:
        resp["error"] = f"model {name} is not served by the B200 runtime"
        return resp
    desc = cat[name][0]()
    ctx = int(request.get("context_length") or 0) or min(desc.max_pos, 8192)
    par = int(request.get("num_parallel") or 0) or DEFAULT_MAX_NUM_SEQS
    try:
        est = memory_estimate(desc, EngineConfig(max_seqs=par, max_ctx=ctx,
                                                 max_batched_tokens=int(request.get("batch_size") or 0) or 16384))
    except HBError as e:
        resp["error"] = str(e)
        return resp
    total = est["weights"] + est["kv"] + est["workspace"]
    resp.update(success=True, architecture="llama" if desc.arch == 0 else "bert", block_count=desc.layers,
                configurations=[{"name": "single_gpu", "gpu_count": 1, "gpu_sizes": [total], "total_memory": total,
                                 "vram_required": total, "weights_memory": est["weights"], "kv_cache": est["kv"],
                                 "graph_memory": est["workspace"], "tensor_split": "", "layers_on_gpu": desc.layers + 1,
                                 "total_layers": desc.layers + 1, "fully_loaded": True}],
                response_time_ms=int((time.time() - t0) * 1000))
    return resp


@dataclass
class B200RuntimeParams:
    model: str
    gpu_index: int = 0                       # CreateRunnerSlotAttributes.gpu_index (types/runner.go:92-104)
    model_memory_requirement: int = 0        # bytes
    per_gpu_memory: int = 0
    context_length: int = 0
    args: List[str] = field(default_factory=list)
    desc: Optional[ModelDesc] = None         # overrides the catalogue (tests)
    state_dict: Optional[dict] = None        # HF-named tensors; None -> random init
    seed: int = 0
    serve_http: bool = True
    tokenizer: Optional[object] = None        # a tokenizer object, or the path of a HF tokenizer.json (loaded natively: hb_tok_*)
    engine_factory: Optional[Callable] = None  # EngineConfig -> engine; default: the CUDA engine (tests inject a stand-in)


class B200Runtime:
    """runner.Runtime for the in-process engine."""

    def __init__(self, params: B200RuntimeParams):
        self.p = params
        self.parsed = parse_vllm_args(params.args)
        self.engine: Optional[Engine] = None
        self.server = None
        self._url = ""
        self._status = ""

    # ---- Runtime interface
    def start(self) -> None:
        """Runtime.Start: create the engine on gpu_index inside the slot's memory budget, load weights, start the
        step loop and the OpenAI-compatible front.  Any failure leaves nothing allocated (Slot.Create's deferred
        Stop, slot.go:113-140)."""
        desc = self.p.desc
        embed = self.parsed.task_embed
        if desc is None:
            if self.p.model not in MODEL_CATALOGUE:
                raise HBError(-5, f"model {self.p.model!r} is not in the B200 runtime catalogue")
            mk, embed = MODEL_CATALOGUE[self.p.model]
            desc = mk()
        max_ctx = self.parsed.max_model_len or self.p.context_length or min(desc.max_pos, 8192)
        cfg = EngineConfig(device=self.p.gpu_index,
                           memory_budget_bytes=memory_budget(self.p.model_memory_requirement, self.p.per_gpu_memory,
                                                             self.parsed.gpu_memory_utilization),
                           max_seqs=self.parsed.max_num_seqs, max_ctx=max_ctx, use_cuda_graphs=1,
                           max_batched_tokens=self.parsed.max_num_batched_tokens or 16384,
                           enable_prefix_cache=int(self.parsed.enable_prefix_caching and not embed),
                           decode_with_prefill=1)  # running streams keep decoding while long prompts are prefilled
        eng = (self.p.engine_factory or Engine)(cfg)
        try:
            if self.p.state_dict is not None:
                eng.load_state_dict(desc, self.p.state_dict)
            else:
                eng.load_random(desc, self.p.seed)
            if desc.arch == configs.LLAMA:
                eng.start()
            self.engine = eng
            self.is_embed = embed or desc.arch == configs.BERT
            if self.p.serve_http:
                from .server import OpenAIServer
                tok = self.p.tokenizer
                if isinstance(tok, (str, bytes)) or hasattr(tok, "__fspath__"):  # path of a HF tokenizer.json: native tokenizer
                    from .tokenizer import NativeTokenizer
                    tok = NativeTokenizer(tok)
                self.server = OpenAIServer(self, tok)
                self._url = self.server.start()
            self._status = "running"
        except Exception:
            eng.close()
            self.engine = None
            raise

    def stop(self) -> None:
        """Runtime.Stop: synchronous; all device memory is released before returning (server.go:801-817 then
        polls nvidia-smi for the memory to come back)."""
        if self.server:
            self.server.stop()
            self.server = None
        if self.engine:
            self.engine.close()
            self.engine = None
        self._status = ""

    def pull_model(self, model: str, progress: Optional[Callable] = None) -> None:
        """Runtime.PullModel: weights are uploaded through hb_model_tensor_set at Start; nothing to pull."""
        if progress:
            progress({"status": "success", "completed": 1, "total": 1})

    def warm(self, model: str) -> None:
        """Runtime.Warm (vllm_runtime.go:391-489 sends "Say the word 'warm'."): one short generation / encode."""
        if self.is_embed:
            self.engine.embed([[1, 2, 3]])
        else:
            rid = self.engine.submit([1, 2, 3, 4], Sampling(max_tokens=2))
            fin = 0
            while not fin:
                self.engine.wait(rid, 10000)
                _, fin = self.engine.poll(rid)
            self.engine.release(rid)

    def list_models(self) -> List[str]:
        return [self.p.model]

    def version(self) -> str:
        return VERSION

    def status(self) -> str:
        """Non-empty == running (scheduler/scheduler.go:940); a sticky CUDA error empties it."""
        if not self.engine:
            return ""
        st = self.engine.stats()
        return "" if st["cuda_error"] else f"running kv_pages_free={st['kv_pages_free']}/{st['kv_pages_total']} " \
                                           f"running={st['running']} waiting={st['waiting']}"

    def runtime(self) -> str:
        return "vllm"  # Option A: unchanged on the wire (types/runner.go:81-86)

    def url(self) -> str:
        return self._url

    def command_line(self) -> str:
        return "helix-b200 (in-process) " + " ".join(self.p.args)
