#!/usr/bin/env python
"""bench.py — BASELINE.json metric: tokens/sec, Llama-3-8B prefill+decode, 32 concurrent sessions x seq 2048
(configs[1]) on N B200s (one engine replica per GPU, weights NCCL-broadcast from rank 0; weak scaling).

One "step" = every session of the workload served once: `sessions` prompts of `seq` random token ids are
submitted through the C ABI (host buffers), prefilled (continuous batching, whole-prompt admission) and each
decodes `decode` tokens over the paged KV cache.  tokens = sessions * (seq + decode).

  value : tokens / device time of the forward passes (CUDA events on the engine stream, inputs resident)
  e2e   : tokens / wall time through hb_submit/hb_wait/hb_poll with host buffers (H2D + D2H inside)
  roofline / kernels : one extra profiled step (CUDA-event span around every launch, same stream)
  cpu_baseline / --impl reference : the oracle port (numpy fp32) on the host cores, bounded sample

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "tokens/sec Llama-3-8B prefill+decode (32 sessions x seq 2048, continuous batching)"
UNIT = "tokens/s"


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            d = json.load(f)
        return {"hbm_gbs": d["hbm_gbs"], "bf16_tflops": d["bf16_tflops"],
                "bf16_tflops_sustained": d.get("bf16_tflops_sustained", d["bf16_tflops"]), "source": "measured"}
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0, "source": "fallback"}


def ncu_traffic(family):
    p = os.path.join(ROOT, "profiles", "ncu_traffic.json")
    if os.path.exists(p):
        with open(p) as f:
            return json.load(f).get(family)
    return None


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md recipe)."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.proc, self.lines = index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "200"], stdout=subprocess.PIPE, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for ln in self.proc.stdout:
            self.lines.append(ln.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, pw, reasons = [], [], [], set()
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0])); mx.append(float(f[1])); pw.append(float(f[2]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        load = [s for s, p in zip(sm, pw) if p > 300] or sm
        return {"sm_mhz": statistics.median(load) if load else None, "sm_max_mhz": max(mx) if mx else None,
                "power_w_max": max(pw) if pw else None, "samples": len(sm), "reasons": sorted(reasons)}


def make_prompts(rank, sessions, seq, vocab):
    return [np.random.default_rng(1000 * rank + i).integers(0, vocab, size=seq, dtype=np.int64).astype(np.int32)
            for i in range(sessions)]


def serve_once(e, hb, prompts, decode):
    """All sessions through the public C-ABI path; returns generated token count."""
    sp = hb.Sampling(max_tokens=decode, temperature=0.0)
    rids = [e.submit(p, sp) for p in prompts]
    got = 0
    for r in rids:
        fin = 0
        while not fin:
            e.wait(r, 60000)
            toks, fin = e.poll(r)
            got += len(toks)
        e.release(r)
    return got


def serve_latency(e, hb, prompts, decode):
    """Untimed extra pass: time to first token and inter-token latency as a client polling the
    C ABI sees them, all sessions submitted at t=0."""
    sp = hb.Sampling(max_tokens=decode, temperature=0.0)
    t0 = time.perf_counter()
    rids = [e.submit(p, sp) for p in prompts]
    stamps = {r: [] for r in rids}
    active = list(rids)
    while active:
        e.wait(active[0], 5)
        now = time.perf_counter()
        for r in list(active):
            toks, fin = e.poll(r)
            stamps[r] += [now] * len(toks)
            if fin:
                active.remove(r)
                e.release(r)
    ttft = np.array([s[0] - t0 for s in stamps.values() if s]) * 1e3
    itl = np.concatenate([np.diff(s) for s in stamps.values() if len(s) > 1]) * 1e3
    q = lambda a, p: float(np.percentile(a, p)) if len(a) else None
    return {"ttft_ms": {"p50": q(ttft, 50), "max": q(ttft, 100)}, "itl_ms": {"p50": q(itl, 50), "p99": q(itl, 99)},
            "note": "all sessions submitted at t=0; prefill has priority, so TTFT includes the queued prompts ahead"}


def bench_bge(args):
    """BASELINE configs[2] (extra line, not the headline): bge-base-en-shaped encoder, batch-encode `--chunks` x 512-token
    synthetic chunks through hb_embed (host token buffers in, fp32 vectors out)."""
    import torch
    import helix_b200 as hb
    from helix_b200 import configs
    desc = configs.bge_base()
    e = hb.Engine(hb.EngineConfig(device=0, max_seqs=64, max_ctx=512, max_batched_tokens=args.max_batched_tokens * 4))
    e.load_random(desc, 2)
    n, L = args.chunks, 512
    rng = np.random.default_rng(2)
    if args.ragged:  # SURVEY.md §8d config 3, varlen variant: chunk lengths U[64, 512]
        lens = rng.integers(64, L + 1, size=n)
    else:
        lens = np.full(n, L)
    offs = np.concatenate([[0], np.cumsum(lens)]).astype(np.int32)
    total_tokens = int(offs[-1])
    toks = rng.integers(0, desc.vocab, size=total_tokens, dtype=np.int64).astype(np.int32)
    out = np.empty((n, desc.hidden), np.float32)
    warm = min(n, 2048)
    for _ in range(max(1, args.warmup)):
        e.embed_flat(toks[:offs[warm]], offs[:warm + 1], out[:warm])
    torch.cuda.synchronize()
    s0 = e.stats()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        e.embed_flat(toks, offs, out)
    torch.cuda.synchronize()
    wall = (time.perf_counter() - t0) / args.steps
    s1 = e.stats()
    dev = (s1["gpu_ms_prefill"] - s0["gpu_ms_prefill"]) / args.steps / 1e3
    e.set_profile(True)
    p0 = e.stats()
    sub = min(n, 8192)
    e.embed_flat(toks[:offs[sub]], offs[:sub + 1], out[:sub])
    p1 = e.stats()
    e.set_profile(False)
    fam = {}
    for i, name in ((0, "gemm"), (1, "attention"), (3, "row_kernels")):
        ms = p1["prof_ms"][i] - p0["prof_ms"][i]
        w = p1["prof_work"][i] - p0["prof_work"][i]
        fam[name] = {"ms": ms, "launches": p1["prof_launches"][i] - p0["prof_launches"][i],
                     "tflops" if i < 2 else "gbs": (w / (ms * 1e-3) / (1e12 if i < 2 else 1e9)) if ms else 0.0}
    # GEMMs 169.9 MFLOP/token + bidirectional attention 4*len*768 FLOP/token
    flops = float(total_tokens) * 169.9e6 + float((lens.astype(np.float64) ** 2).sum()) * 4 * 768
    peaks = measured_peaks()
    line = {"metric": "chunks/sec bge-base-en-shaped batch encode (512-token chunks)", "value": n / dev, "unit": "chunks/s",
            "n_gpus": 1, "steps": args.steps, "warmup": args.warmup, "ms_per_step": dev * 1e3, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
            "config": {"workload": f"bge-base-en shape random-init, {n} chunks x " + ("U[64,512]" if args.ragged else "512") +
                       " tokens, CLS+L2, fp32 out"},
            "e2e": {"value": n / wall, "unit": "chunks/s", "h2d_bytes_per_step": int(toks.nbytes * 2), "d2h_bytes_per_step": int(out.nbytes)},
            "model_tflops": flops / dev / 1e12, "frac_of_measured_sustained": flops / dev / 1e12 / peaks["bf16_tflops_sustained"],
            "gpu_launches": int(s1["kernel_launches"] - s0["kernel_launches"]), "tokens_per_s": total_tokens / dev,
            "kernels_profiled_subset": fam, "finite": bool(np.isfinite(out).all()), "unit_norm_err": float(np.abs(np.linalg.norm(out, axis=1) - 1).max())}
    print(json.dumps(line), flush=True)
    e.close()


def cpu_sample(layers=4, prompt=512, decode=8, threads=None):
    """Oracle port on the host cores: L8B shape truncated to `layers` layers, 1 session; linear extrapolation to 32."""
    from helix_b200 import configs
    from oracle.llama_ref import LlamaOracle
    d = configs.llama3_8b()
    d.layers = layers
    rng = np.random.default_rng(0)
    block = (rng.standard_normal(1 << 22, dtype=np.float32) * 0.02)

    def filled(shape):
        n = int(np.prod(shape))
        return np.resize(block, n).reshape(shape)
    H, F, V, D = d.hidden, d.ffn, d.vocab, d.head_dim
    sd = {"model.embed_tokens.weight": filled((V, H)), "lm_head.weight": filled((V, H)), "model.norm.weight": np.ones(H, np.float32)}
    for i in range(layers):
        p = f"model.layers.{i}."
        sd[p + "input_layernorm.weight"] = np.ones(H, np.float32)
        sd[p + "post_attention_layernorm.weight"] = np.ones(H, np.float32)
        sd[p + "self_attn.q_proj.weight"] = filled((d.heads * D, H))
        sd[p + "self_attn.k_proj.weight"] = filled((d.kv_heads * D, H))
        sd[p + "self_attn.v_proj.weight"] = filled((d.kv_heads * D, H))
        sd[p + "self_attn.o_proj.weight"] = filled((H, d.heads * D))
        sd[p + "mlp.gate_proj.weight"] = filled((F, H))
        sd[p + "mlp.up_proj.weight"] = filled((F, H))
        sd[p + "mlp.down_proj.weight"] = filled((H, F))
    o = LlamaOracle(d, sd)
    toks = rng.integers(0, V, size=prompt).astype(np.int32)

    def run():
        t0 = time.perf_counter()
        o.greedy(toks, decode)
        return time.perf_counter() - t0
    return run, (prompt + decode), 32.0 / layers, f"oracle port (numpy fp32), L8B shape truncated to {layers}/32 layers " \
        f"(time x{32 // layers}), 1 session x {prompt} prompt + {decode} decode tokens; stand-in for the reference's " \
        "llama.cpp CPU path, which cannot run offline"


def reference_arm(args, rank, world):
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    run, toks, scale, sample = cpu_sample()
    for _ in range(min(args.warmup, 1)):
        run()
    ts = [run() for _ in range(max(1, min(args.steps, 3)))]
    t = statistics.mean(ts) * scale
    v = toks / t
    line = {"impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": len(ts),
            "warmup": min(args.warmup, 1), "ms_per_step": t * 1e3, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": "Llama-3-8B-shaped random-init, 32 sessions x seq 2048 + decode (bounded CPU sample)"},
            "cpu_baseline": {"value": v, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="native", choices=["native", "reference"])
    ap.add_argument("--sessions", type=int, default=32)
    ap.add_argument("--seq", type=int, default=2048)
    ap.add_argument("--decode", type=int, default=128)
    ap.add_argument("--max-batched-tokens", type=int, default=16384)
    ap.add_argument("--no-graphs", action="store_true")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--layers", type=int, default=0, help="debug: truncate the model (INVALID as a bench number)")
    ap.add_argument("--workload", default="llama8b", choices=["llama8b", "bge"])
    ap.add_argument("--chunks", type=int, default=100000)
    ap.add_argument("--ragged", action="store_true", help="bge workload: chunk lengths U[64,512] instead of 512")
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.impl == "reference":
        reference_arm(args, rank, world)
        return
    if args.workload == "bge":
        if rank == 0:
            bench_bge(args)
        return

    import torch
    import torch.distributed as dist
    import helix_b200 as hb
    from helix_b200 import configs
    from helix_b200.replica import ArenaView, broadcast_buffer

    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    desc = configs.llama3_8b()
    if args.layers:
        desc.layers = args.layers
    max_ctx = ((args.seq + args.decode + 63) // 64 + 1) * 64
    cfg = hb.EngineConfig(device=local_rank, max_seqs=args.sessions, max_ctx=max_ctx,
                          max_batched_tokens=args.max_batched_tokens, use_cuda_graphs=0 if args.no_graphs else 1)
    e = hb.Engine(cfg)
    e.load_random(desc, seed=0 if rank == 0 else 7919 * rank)  # replicas start different ...
    bcast = None
    if world > 1:  # ... and receive rank 0's weights with ONE NCCL broadcast of the arena
        ptr, nbytes = e.weights_arena()
        arena = torch.as_tensor(ArenaView(ptr, nbytes), device=f"cuda:{local_rank}")
        probe = arena[:: max(1, nbytes // 65536)].clone()
        dist.broadcast(torch.zeros(1 << 18, device=f"cuda:{local_rank}"), src=0)  # NCCL channel setup is not part of the measured copy
        barrier()
        t0 = time.perf_counter()
        broadcast_buffer(dist, arena, src=0)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        chk = arena[:: max(1, nbytes // 65536)].to(torch.int32).sum().reshape(1).float()
        lo, hi = chk.clone(), chk.clone()
        dist.all_reduce(lo, op=dist.ReduceOp.MIN)
        dist.all_reduce(hi, op=dist.ReduceOp.MAX)
        assert float(lo) == float(hi), "weight broadcast mismatch across replicas"
        if rank != 0:
            assert not torch.equal(probe, arena[:: max(1, nbytes // 65536)]), "broadcast did not overwrite replica weights"
        bcast = {"bytes": nbytes, "seconds": dt, "gbs": nbytes / dt / 1e9}

    prompts = make_prompts(rank, args.sessions, args.seq, desc.vocab)
    tokens_per_step = args.sessions * (args.seq + args.decode)
    e.start()
    for _ in range(args.warmup):
        got = serve_once(e, hb, prompts, args.decode)
        assert got == args.sessions * args.decode

    sampler = ClockSampler(local_rank) if rank == 0 else None
    barrier()
    if sampler:
        sampler.start()
    s0 = e.stats()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        serve_once(e, hb, prompts, args.decode)
    torch.cuda.synchronize()
    wall = time.perf_counter() - t0
    s1 = e.stats()
    barrier()
    clocks = sampler.stop() if sampler else None
    dev_ms = (s1["gpu_ms_prefill"] - s0["gpu_ms_prefill"]) + (s1["gpu_ms_decode"] - s0["gpu_ms_decode"])
    times = torch.tensor([dev_ms / 1e3, wall], device="cuda", dtype=torch.float64)
    if world > 1:
        dist.all_reduce(times, op=dist.ReduceOp.MAX)  # max over ranks
    dev_s, wall_s = float(times[0]), float(times[1])
    launches = s1["kernel_launches"] - s0["kernel_launches"]

    latency = serve_latency(e, hb, prompts, args.decode)
    # ---- one profiled step: CUDA-event span around every launch on the engine stream
    e.set_profile(True)
    p0 = e.stats()
    serve_once(e, hb, prompts, args.decode)
    p1 = e.stats()
    e.set_profile(False)
    e.stop()
    fam = ["gemm_prefill", "attn_prefill", "attn_decode", "row_kernels", "gemm_decode"]
    prof = {}
    for i, name in enumerate(fam):
        ms = p1["prof_ms"][i] - p0["prof_ms"][i]
        work = p1["prof_work"][i] - p0["prof_work"][i]
        n = p1["prof_launches"][i] - p0["prof_launches"][i]
        prof[name] = {"ms": ms, "work": work, "launches": n}
    total_prof_ms = sum(v["ms"] for v in prof.values()) or 1.0
    peaks = measured_peaks()
    g = prof["gemm_prefill"]
    gemm_tf = g["work"] / (g["ms"] * 1e-3) / 1e12 if g["ms"] else 0.0
    gd = prof["gemm_decode"]
    gemm_dec_gbs = gd["work"] / (gd["ms"] * 1e-3) / 1e9 if gd["ms"] else 0.0
    ad = prof["attn_decode"]
    ap_ = prof["attn_prefill"]
    roofline = {"kernel": "gemm_tn_kernel (tcgen05, prefill steps)", "bound": "tensor", "achieved": gemm_tf,
                "peak": peaks["bf16_tflops_sustained"], "unit": "TFLOP/s",
                "frac": gemm_tf / peaks["bf16_tflops_sustained"], "peak_source": peaks["source"] + " sustained cuBLAS bf16",
                "frac_of_nominal_2250": gemm_tf / 2250.0, "launches": g["launches"],
                "avg_launch_ms": g["ms"] / g["launches"] if g["launches"] else None,
                "share_of_step": g["ms"] / total_prof_ms, "traffic": ncu_traffic("gemm_prefill")}
    kernels = {
        "gemm_decode": {"bound": "hbm", "achieved": gemm_dec_gbs, "peak": peaks["hbm_gbs"], "unit": "GB/s",
                        "frac": gemm_dec_gbs / peaks["hbm_gbs"], "share_of_step": gd["ms"] / total_prof_ms,
                        "launches": gd["launches"], "traffic": ncu_traffic("gemm_decode")},
        "attn_decode": {"bound": "hbm", "achieved": ad["work"] / (ad["ms"] * 1e-3) / 1e9 if ad["ms"] else 0.0,
                        "peak": peaks["hbm_gbs"], "unit": "GB/s", "share_of_step": ad["ms"] / total_prof_ms,
                        "launches": ad["launches"]},
        "attn_prefill": {"bound": "tensor", "achieved": ap_["work"] / (ap_["ms"] * 1e-3) / 1e12 if ap_["ms"] else 0.0,
                         "peak": peaks["bf16_tflops_sustained"], "unit": "TFLOP/s",
                         "share_of_step": ap_["ms"] / total_prof_ms, "launches": ap_["launches"]},
        "row_kernels": {"bound": "hbm", "share_of_step": prof["row_kernels"]["ms"] / total_prof_ms,
                        "launches": prof["row_kernels"]["launches"]},
    }
    for k in ("attn_decode", "attn_prefill"):
        kernels[k]["frac"] = kernels[k]["achieved"] / kernels[k]["peak"]
    # phase view (device time of the timed region)
    pre_ms = (s1["gpu_ms_prefill"] - s0["gpu_ms_prefill"]) / args.steps
    dec_ms = (s1["gpu_ms_decode"] - s0["gpu_ms_decode"]) / args.steps
    phases = {"prefill_tokens_per_s": args.sessions * args.seq / (pre_ms * 1e-3) if pre_ms else None,
              "decode_tokens_per_s": args.sessions * args.decode / (dec_ms * 1e-3) if dec_ms else None,
              "prefill_ms": pre_ms, "decode_ms": dec_ms,
              "prefill_model_tflops": (args.sessions * args.seq * 14.50e9 + args.sessions * 1.05e9) / (pre_ms * 1e-3) / 1e12
              if pre_ms and not args.layers else None}

    if rank == 0:
        # per-step host<->device traffic of the public path: prompt ids + per-step metadata in, sampled ids out
        h2d = args.sessions * args.seq * 4 * 3 + args.decode * args.sessions * 4 * 8
        d2h = args.sessions * (args.decode) * 4
        line = {"metric": METRIC, "value": world * tokens_per_step * args.steps / dev_s, "unit": UNIT, "n_gpus": world,
                "steps": args.steps, "warmup": args.warmup, "ms_per_step": dev_s * 1e3 / args.steps,
                "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
                "config": {"workload": f"Llama-3-8B random-init bf16, {args.sessions} sessions/GPU x ({args.seq} prompt + "
                                       f"{args.decode} decode) tokens, continuous batching, paged KV (page 64), "
                                       f"prefill budget {args.max_batched_tokens} tokens/step, CUDA-graph decode="
                                       f"{not args.no_graphs}",
                           "parallelism": f"replica x{world} (NCCL weight broadcast, no data-path collective)",
                           "l2": "weights 16 GB + KV >> 126 MB L2: inputs larger than L2, no flush needed",
                           "value_timing": "CUDA events on the engine stream around every forward pass, summed; max over ranks",
                           "broadcast": bcast, "layers": desc.layers},
                "e2e": {"value": world * tokens_per_step * args.steps / wall_s, "unit": UNIT, "ms_per_step": wall_s * 1e3 / args.steps,
                        "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                        "api": "hb_submit/hb_wait/hb_poll (C ABI, host buffers, step-loop thread)"},
                "gpu_launches": int(launches), "clocks": clocks, "roofline": roofline, "kernels": kernels, "phases": phases,
                "latency": latency}
        if world == 1 and not args.no_cpu_baseline:
            run, toks, scale, sample = cpu_sample()
            t = run() * scale
            line["cpu_baseline"] = {"value": toks / t, "unit": UNIT, "cores": os.cpu_count(), "kind": "port", "sample": sample}
        print(json.dumps(line), flush=True)
    e.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
