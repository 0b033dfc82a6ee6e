"""GPU: request fields and runtime contracts added in ABI 2, through the C ABI, each against its oracle:
presence/frequency penalties and log-probabilities (oracle/sampling_ref.py), the decoder embedder (`--task embed` on a
Llama-architecture model, oracle LlamaOracle.embed), replica loading by one NCCL broadcast inside the library
(hb_model_load_broadcast), Stop() returning every byte of device memory, and the plain-C host (tests/abi_host.c) that
drives the whole boundary without Python — what the Go shim's cgo calls compile down to."""
import ctypes as C
import os
import subprocess
import threading

import numpy as np
import pytest

import helix_b200 as hb
from helix_b200 import configs
from helix_b200.engine import CAPTURE_STEP_LOGITS, replica_unique_id
from oracle import sampling_ref as S
from oracle import weights
from oracle.llama_ref import LlamaOracle

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def tol(ref):
    return 2e-2 * max(1.0, float(np.abs(ref).max()))


@pytest.mark.parametrize("graphs", [0, 1])
def test_penalties_and_logprobs_through_the_engine(graphs):
    d = configs.tiny_llama(layers=2, head_dim=64, vocab=1000)
    sd = weights.llama_state_dict(d, 21, 0.05)
    prompt = weights.random_tokens(22, 40, d.vocab)
    pres, freq, W, n = 0.7, 1.2, 5, 24
    with hb.Engine(hb.EngineConfig(max_seqs=4, max_ctx=256, max_batched_tokens=256, use_cuda_graphs=graphs)) as e:
        e.load_state_dict(d, sd)
        sp = hb.Sampling(max_tokens=n, capture=CAPTURE_STEP_LOGITS, logprobs=W, presence_penalty=pres, frequency_penalty=freq)
        plain = hb.Sampling(max_tokens=n)
        rids, outs = e.generate([prompt, prompt], sp)           # two identical penalised rows in one batch
        rp, op = e.generate([prompt], plain)
        rows = e.captured_logits(rids[0], CAPTURE_STEP_LOGITS)   # logits as sampled from: AFTER the penalties
        ids, lps = e.logprobs(rids[0])
        ids2, lps2 = e.logprobs(rids[0], first_row=3, max_rows=2)
        with pytest.raises(hb.HBError):
            e.submit(prompt, hb.Sampling(max_tokens=4, logprobs=22))
    toks = outs[0]
    assert outs[1] == toks and len(toks) == n
    assert toks != op[0]                                          # the penalties changed the greedy continuation
    assert ids.shape == (n, W) and np.array_equal(ids2, ids[3:5]) and np.array_equal(lps2, lps[3:5])
    o = LlamaOracle(d, sd)
    logits = o.forward(prompt)[-1]
    for i, t in enumerate(toks):
        want = S.penalised(logits, toks[:i], pres, freq)
        b = tol(want)
        assert np.abs(rows[i] - want).max() <= b, i
        best = int(np.argmax(want))
        assert t == best or want[best] - want[t] <= 2 * b, (i, t, best)
        wi, wl = S.logprob_record(rows[i], t, W)                  # same logits the kernel saw: ids exact, values tight
        assert ids[i].tolist() == wi.tolist(), i
        assert np.abs(lps[i] - wl).max() < 1e-3
        _, ol = S.logprob_record(want, t, 1)                      # and against the oracle's own distribution
        assert abs(lps[i, 0] - ol[0]) <= 2 * b
        logits = o.forward([t])[-1]


def test_decoder_embedder_last_token_pooling_vs_oracle():
    """hb_embed on a Llama-architecture engine: causal pass without touching the KV pool, last-token pooling of the
    final-norm hidden state, L2 — the reference's default embedding model (MrLight/dse-qwen2-2b-mrl-v1, --task embed) is a
    decoder.  max-abs <= 1e-2 and cosine >= 0.9999 vs the fp32 oracle; batch composition must not matter."""
    d = configs.tiny_llama(layers=3, head_dim=128, vocab=1000, rope_scaling=True)
    sd = weights.llama_state_dict(d, 31, 0.05)
    seqs = [weights.random_tokens(40 + i, n, d.vocab) for i, n in enumerate([1, 5, 64, 129, 300, 17])]
    with hb.Engine(hb.EngineConfig(max_seqs=8, max_ctx=512, max_batched_tokens=512)) as e:
        e.load_state_dict(d, sd)
        e.start()
        rid = e.submit(seqs[2], hb.Sampling(max_tokens=8))        # a generation running next to the embed calls
        got = e.embed(seqs)
        solo = e.embed([seqs[3]])
        while not e.poll(rid)[1]:
            e.wait(rid, 1000)
        e.release(rid)
        st = e.stats()
    o = LlamaOracle(d, sd)
    ref = np.stack([o.embed(s) for s in seqs])
    assert np.abs(got - ref).max() <= 1e-2 and float((got * ref).sum(-1).min()) >= 0.9999
    assert np.array_equal(solo[0], got[3])
    assert st["kv_pages_free"] == st["kv_pages_total"]            # embeds never take pages


def test_qwen2_style_decoder_generation_and_embedding(golden_dir):
    """Qwen2-family decoder (hb_model_desc.qkv_bias: q/k/v biases; GQA group 6; tied embeddings) through the engine: prompt
    logits and greedy decode (bias in the prefill GEMM epilogue and in the decode RoPE kernel) vs the HF fixture and the
    oracle, and `--task embed` pooling vs HF's last-token hidden state."""
    g = np.load(os.path.join(golden_dir, "qwen2_tiny.npz"))
    d = configs.tiny_qwen2(layers=2, vocab=1000)
    sd = weights.llama_state_dict(d, int(g["seed"]), float(g["std"]))
    prompt, n_dec = g["prompt"], len(g["greedy_tokens"])
    from helix_b200.engine import CAPTURE_PROMPT_LOGITS
    with hb.Engine(hb.EngineConfig(max_seqs=4, max_ctx=256, max_batched_tokens=256, use_cuda_graphs=1)) as e:
        e.load_state_dict(d, sd)
        rids, outs = e.generate([prompt], hb.Sampling(max_tokens=n_dec, capture=CAPTURE_PROMPT_LOGITS | CAPTURE_STEP_LOGITS))
        pl = e.captured_logits(rids[0], CAPTURE_PROMPT_LOGITS)
        sl = e.captured_logits(rids[0], CAPTURE_STEP_LOGITS)
        emb = e.embed([prompt, prompt[:5]])
    assert np.abs(pl - g["prompt_logits"]).max() <= tol(g["prompt_logits"])
    o = LlamaOracle(d, sd)
    logits = o.forward(prompt)[-1]
    worst = 0.0
    for i, t in enumerate(outs[0]):
        # 768-wide rows, N(0, 0.05) weights and N(0, 0.1) biases: measured worst 2.2e-2 of |row|_inf (decode step 8), so this
        # case states 3e-2 like the BASELINE-shape tests (tests/test_baseline_shapes_gpu.py explains the scaling)
        bound = 1.5 * tol(logits)
        worst = max(worst, float(np.abs(sl[i] - logits).max()) / bound)
        assert np.abs(sl[i] - logits).max() <= bound, i
        best = int(np.argmax(logits))
        assert t == best or logits[best] - logits[t] <= 2 * bound
        logits = o.forward([t])[-1]
    print(f"\n[qwen2 tiny] worst |dlogit|/bound = {worst:.3f} (bound 3e-2*max(1,|row|_inf))")
    assert outs[0][:4] == g["greedy_tokens"].tolist()[:4]
    err, cos = float(np.abs(emb[0] - g["embedding"]).max()), float(emb[0] @ g["embedding"])
    print(f"[qwen2 tiny] embedding vs HF: max|d| {err:.2e}, cosine {cos:.6f}")
    # N(0, 0.05) weights / N(0, 0.1) biases make this the noisiest configuration in the suite (logit error 2.2e-2 above):
    # cosine is held to 0.9995 here, 0.9999 on every other embedding test
    assert err <= 1e-2 and cos >= 0.9995, (err, cos)
    assert np.abs(emb[1] - o.embed(prompt[:5])).max() <= 1e-2


@pytest.mark.parametrize("arch", ["llama", "qwen2"])
def test_gguf_checkpoint_loads_and_generates_like_the_oracle(tmp_path, arch):
    """hb_model_load_gguf: a llama.cpp-format file with the tensor types Ollama's blobs use (Q4_0 / Q4_K / Q5_K / Q6_K / Q8_0 /
    F32 norms; q/k rows in llama.cpp's permuted order) is dequantised into the arena; greedy generation then matches the
    fp32 oracle run on the (bf16-rounded) dequantised weights."""
    import sys
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from gguf_ref import llama_gguf
    d = configs.tiny_llama(layers=2, head_dim=64, vocab=1000) if arch == "llama" else configs.tiny_qwen2(layers=2, vocab=1000)
    if arch == "llama":
        d.hidden, d.heads, d.kv_heads, d.ffn = 512, 8, 2, 1024     # rows of 512 / 1024: whole K-quant super-blocks
    rng = np.random.default_rng(11)
    path = tmp_path / f"{arch}.gguf"
    sd = llama_gguf(path, d, rng, {"embd": "Q8_0", "attn": "Q4_0", "v": "Q6_K", "ffn": "Q4_K", "down": "Q5_K", "output": "Q6_K"}, arch=arch)
    sd = {k: weights.to_bf16_f32(v) for k, v in sd.items()}   # what the arena holds: one bf16 rounding of the dequantised value
    prompt = weights.random_tokens(12, 40, d.vocab)
    with hb.Engine(hb.EngineConfig(max_seqs=2, max_ctx=256, max_batched_tokens=256, use_cuda_graphs=1)) as e:
        e.load_gguf(path)
        assert e.desc.hidden == d.hidden and e.desc.qkv_bias == d.qkv_bias
        rids, outs = e.generate([prompt], hb.Sampling(max_tokens=8, capture=CAPTURE_STEP_LOGITS))
        sl = e.captured_logits(rids[0], CAPTURE_STEP_LOGITS)
    o = LlamaOracle(e.desc, sd)
    logits = o.forward(prompt)[-1]
    for i, t in enumerate(outs[0]):
        bound = 1.5 * tol(logits)
        assert np.abs(sl[i] - logits).max() <= bound, (i, float(np.abs(sl[i] - logits).max()), bound)
        best = int(np.argmax(logits))
        assert t == best or logits[best] - logits[t] <= 2 * bound
        logits = o.forward([t])[-1]


def _free_bytes():
    import torch
    return torch.cuda.mem_get_info(0)[0]


def test_stop_returns_all_device_memory():
    """Runtime.Stop contract (the runner polls nvidia-smi for the memory to come back,
    api/pkg/runner/server.go:801-817): after hb_engine_destroy the device's free memory is back at its level from before
    hb_engine_create — with a stream handler still parked in hb_wait when the engine goes away."""
    import torch
    torch.cuda.init()
    torch.cuda.synchronize()
    d = configs.tiny_llama(layers=2, head_dim=64, vocab=1000)
    warm = hb.Engine(hb.EngineConfig(max_seqs=2, max_ctx=128, max_batched_tokens=128))  # context / module load is one-off
    warm.load_random(d, 1)
    warm.close()
    before = _free_bytes()
    e = hb.Engine(hb.EngineConfig(max_seqs=16, max_ctx=4096, max_batched_tokens=4096, use_cuda_graphs=1,
                                  memory_budget_bytes=6 << 30))
    e.load_random(d, 1)
    e.start()
    rid = e.submit(weights.random_tokens(1, 100, d.vocab), hb.Sampling(max_tokens=3000))
    during = _free_bytes()
    assert before - during > (32 << 20)                          # the KV pool (64 MB at this tiny shape) really was allocated
    res = {}
    t = threading.Thread(target=lambda: res.setdefault("rc", e._l.hb_wait(e._h, rid, 60000)))
    for _ in range(50):
        e.wait(rid, 100)
        if e.poll(rid)[0]:
            break
    # drain what is there, then park a waiter on the still-running request and destroy the engine under it
    t.start()
    e.close()
    t.join(10)
    assert not t.is_alive()
    after = _free_bytes()
    assert abs(after - before) <= (8 << 20), (before, during, after)   # allocator granularity only


def test_replica_broadcast_inside_the_library_single_rank():
    """hb_replica_unique_id + hb_model_load_broadcast with world = 1: the NCCL path is exercised end to end on one GPU
    (communicator, broadcast of the arena, no torch.distributed anywhere)."""
    d = configs.tiny_llama(layers=2, head_dim=64, vocab=1000)
    uid = replica_unique_id()
    assert len(uid) == 128 and any(uid)
    with hb.Engine(hb.EngineConfig(max_seqs=2, max_ctx=128, max_batched_tokens=128)) as e:
        e.load_random(d, 3)
        _, a = e.generate([[1, 2, 3]], hb.Sampling(max_tokens=4))
        sec = e.load_broadcast(d, uid, 0, 1)
        _, b = e.generate([[1, 2, 3]], hb.Sampling(max_tokens=4))
        assert sec > 0 and a == b
        with pytest.raises(hb.HBError):
            e.load_broadcast(d, uid, 1, 1)


def test_replica_broadcast_two_gpus():
    """Two engines on two devices of one box, one thread each (how a Go runner would hold its per-GPU runtimes): rank 1
    starts empty, receives rank 0's arena by ncclBroadcast over NVLink and then generates exactly rank 0's tokens."""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs (gpurun --gpus 2)")
    d = configs.tiny_llama(layers=2, head_dim=64, vocab=1000)
    uid = replica_unique_id()
    engines = [hb.Engine(hb.EngineConfig(device=i, max_seqs=2, max_ctx=128, max_batched_tokens=128)) for i in range(2)]
    engines[0].load_random(d, 5)
    errs = []

    def run(rank):
        try:
            engines[rank].load_broadcast(d, uid, rank, 2)
        except Exception as ex:  # noqa: BLE001
            errs.append(ex)
    ts = [threading.Thread(target=run, args=(r,)) for r in range(2)]
    [t.start() for t in ts]
    [t.join(120) for t in ts]
    assert not errs, errs
    outs = [e.generate([[7, 8, 9, 10]], hb.Sampling(max_tokens=6))[1] for e in engines]
    [e.close() for e in engines]
    assert outs[0] == outs[1]


def test_plain_c_host_drives_the_whole_boundary():
    """tests/abi_host.c is compiled by build() with gcc against include/helix_b200.h ONLY and linked to libhelixb200.so:
    create -> load (begin / tensor_set / finish) -> start -> submit / wait / poll -> logprobs -> embed -> replica id ->
    stats -> destroy, printing token ids this test compares with the Python binding's for the same seeded weights."""
    exe = os.path.join(ROOT, "tests", "abi_host")
    if not os.path.exists(exe):
        pytest.fail("tests/abi_host not built: run __graft_entry__.build()")
    env = dict(os.environ, LD_LIBRARY_PATH=os.path.join(ROOT, "helix_b200") + ":" + os.environ.get("LD_LIBRARY_PATH", ""))
    r = subprocess.run([exe], capture_output=True, text=True, timeout=300, env=env)
    assert r.returncode == 0, r.stdout + r.stderr
    lines = dict(l.split(":", 1) for l in r.stdout.splitlines() if ":" in l)
    c_tokens = [int(x) for x in lines["tokens"].split()]
    # same deterministic weights through the Python binding
    d = configs.tiny_llama(layers=2, head_dim=64, vocab=1000)
    with hb.Engine(hb.EngineConfig(max_seqs=4, max_ctx=256, max_batched_tokens=256)) as e:
        e.load_state_dict(d, host_weights(d))
        _, outs = e.generate([list(range(1, 33))], hb.Sampling(max_tokens=8))
    assert c_tokens == outs[0]
    assert lines["embed_norm"].strip().startswith("1.000") and lines["status"].strip() == "ok"


def host_weights(d):
    """The deterministic pattern tests/abi_host.c fills its tensors with: bf16 of ((i * 2654435761 + salt) mod 2001 - 1000) / 25000,
    norm gains 1."""
    H, F, V, D = d.hidden, d.ffn, d.vocab, d.head_dim
    names = [("model.embed_tokens.weight", (V, H))]
    for i in range(d.layers):
        p = f"model.layers.{i}."
        names += [(p + "input_layernorm.weight", (H,)), (p + "self_attn.q_proj.weight", (d.heads * D, H)),
                  (p + "self_attn.k_proj.weight", (d.kv_heads * D, H)), (p + "self_attn.v_proj.weight", (d.kv_heads * D, H)),
                  (p + "self_attn.o_proj.weight", (H, d.heads * D)), (p + "post_attention_layernorm.weight", (H,)),
                  (p + "mlp.gate_proj.weight", (F, H)), (p + "mlp.up_proj.weight", (F, H)), (p + "mlp.down_proj.weight", (H, F))]
    names += [("model.norm.weight", (H,)), ("lm_head.weight", (V, H))]
    sd = {}
    for salt, (name, shape) in enumerate(names):
        n = int(np.prod(shape))
        if len(shape) == 1:
            sd[name] = np.ones(shape, np.float32)
        else:
            i = np.arange(n, dtype=np.uint64)
            v = ((i * np.uint64(2654435761) + np.uint64(salt * 7919)) % np.uint64(2001)).astype(np.float32)
            sd[name] = weights.to_bf16_f32(((v - 1000.0) / 25000.0).reshape(shape))
    return sd
