"""Generates tests/golden/*.npz with HF transformers (fp32, CPU) — run in the build container:

    python tests/golden/gen_golden.py

HF transformers 5.5.0 `LlamaForCausalLM` / `BertModel` are the pinned comparators of the oracle
(the reference itself holds no golden vector for this path and cannot run offline).
Weights come from oracle/weights.py (seeded, bf16-representable) so nothing but seeds is stored.
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from helix_b200 import configs  # noqa: E402
from oracle import weights  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))
torch.manual_seed(0)
torch.set_grad_enabled(False)


def hf_llama(d, sd):
    from transformers import LlamaConfig, LlamaForCausalLM
    rope = {"rope_theta": d.rope_theta, "rope_type": "default"}
    if d.rope_factor > 0:
        rope = {"rope_theta": d.rope_theta, "rope_type": "llama3", "factor": d.rope_factor,
                "low_freq_factor": d.rope_low_freq_factor, "high_freq_factor": d.rope_high_freq_factor,
                "original_max_position_embeddings": d.rope_orig_max_pos}
    cfg = LlamaConfig(vocab_size=d.vocab, hidden_size=d.hidden, intermediate_size=d.ffn, num_hidden_layers=d.layers,
                      num_attention_heads=d.heads, num_key_value_heads=d.kv_heads, head_dim=d.head_dim,
                      max_position_embeddings=d.max_pos, rms_norm_eps=d.norm_eps, tie_word_embeddings=bool(d.tie_embeddings),
                      rope_parameters=rope, attention_bias=False, mlp_bias=False, attn_implementation="eager")
    m = LlamaForCausalLM(cfg).float().eval()
    missing, unexpected = m.load_state_dict({k: torch.from_numpy(v) for k, v in sd.items()}, strict=False)
    assert not unexpected, unexpected
    assert all("rotary" in k or (d.tie_embeddings and k == "lm_head.weight") for k in missing), missing
    return m


def hf_qwen2(d, sd):
    from transformers import Qwen2Config, Qwen2ForCausalLM
    cfg = Qwen2Config(vocab_size=d.vocab, hidden_size=d.hidden, intermediate_size=d.ffn, num_hidden_layers=d.layers,
                      num_attention_heads=d.heads, num_key_value_heads=d.kv_heads, max_position_embeddings=d.max_pos,
                      rms_norm_eps=d.norm_eps, tie_word_embeddings=bool(d.tie_embeddings), use_sliding_window=False,
                      rope_parameters={"rope_theta": d.rope_theta, "rope_type": "default"}, attn_implementation="eager")
    assert cfg.hidden_size // cfg.num_attention_heads == d.head_dim
    m = Qwen2ForCausalLM(cfg).float().eval()
    missing, unexpected = m.load_state_dict({k: torch.from_numpy(v) for k, v in sd.items()}, strict=False)
    assert not unexpected, unexpected
    assert all("rotary" in k or (d.tie_embeddings and k == "lm_head.weight") for k in missing), missing
    return m


def gen_qwen2(name, d, seed, n_prompt, n_decode, std=0.05):
    """Qwen2-family decoder (q/k/v biases, GQA group 6, tied embeddings): prompt logits, greedy decode and the last-token
    embedding (final-norm hidden state of the last position, L2-normalised: the `--task embed` pooling)."""
    sd = weights.llama_state_dict(d, seed, std)
    m = hf_qwen2(d, sd)
    prompt = weights.random_tokens(seed + 1, n_prompt, d.vocab)
    ids = torch.from_numpy(prompt.astype(np.int64))[None]
    out = m(ids, use_cache=True, output_hidden_states=True)
    prompt_logits = out.logits[0].numpy()
    h = out.hidden_states[-1][0, -1]          # HF applies the final norm before appending the last hidden state
    emb = (h / h.norm().clamp_min(1e-12)).numpy()
    past, logits = out.past_key_values, out.logits[0, -1]
    toks, rows = [], []
    for _ in range(n_decode):
        rows.append(logits.numpy().copy())
        t = int(torch.argmax(logits))
        toks.append(t)
        o = m(torch.tensor([[t]]), past_key_values=past, use_cache=True)
        past, logits = o.past_key_values, o.logits[0, -1]
    np.savez_compressed(os.path.join(OUT, name + ".npz"), seed=seed, std=std, prompt=prompt,
                        prompt_logits=prompt_logits.astype(np.float32), greedy_tokens=np.array(toks, np.int32),
                        step_logits=np.stack(rows).astype(np.float32), embedding=emb.astype(np.float32))
    print(name, "greedy", toks[:8])


def hf_bert(d, sd):
    from transformers import BertConfig, BertModel
    cfg = BertConfig(vocab_size=d.vocab, hidden_size=d.hidden, num_hidden_layers=d.layers, num_attention_heads=d.heads,
                     intermediate_size=d.ffn, hidden_act="gelu", hidden_dropout_prob=0.0, attention_probs_dropout_prob=0.0,
                     max_position_embeddings=d.max_pos, type_vocab_size=d.type_vocab, layer_norm_eps=d.norm_eps,
                     attn_implementation="eager")
    m = BertModel(cfg, add_pooling_layer=False).float().eval()
    missing, unexpected = m.load_state_dict({k: torch.from_numpy(v) for k, v in sd.items()}, strict=False)
    assert not unexpected, unexpected
    assert all("position_ids" in k or "token_type_ids" in k for k in missing), missing
    return m


def gen_llama(name, d, seed, n_prompt, n_decode, std=0.02):
    sd = weights.llama_state_dict(d, seed, std)
    m = hf_llama(d, sd)
    prompt = weights.random_tokens(seed + 1, n_prompt, d.vocab)
    ids = torch.from_numpy(prompt.astype(np.int64))[None]
    prompt_logits = m(ids).logits[0].numpy()
    # greedy decode through HF's own KV cache
    out = m(ids, use_cache=True)
    past = out.past_key_values
    logits = out.logits[0, -1]
    toks, rows = [], []
    for _ in range(n_decode):
        rows.append(logits.numpy().copy())
        t = int(torch.argmax(logits))
        toks.append(t)
        o = m(torch.tensor([[t]]), past_key_values=past, use_cache=True)
        past = o.past_key_values
        logits = o.logits[0, -1]
    np.savez_compressed(os.path.join(OUT, name + ".npz"), seed=seed, std=std, prompt=prompt,
                        prompt_logits=prompt_logits.astype(np.float32), greedy_tokens=np.array(toks, np.int32),
                        step_logits=np.stack(rows).astype(np.float32))
    print(name, "prompt_logits", prompt_logits.shape, "greedy", toks[:8])


def gen_bert(name, d, seed, lens, std=0.05):
    sd = weights.bert_state_dict(d, seed, std)
    m = hf_bert(d, sd)
    seqs = [weights.random_tokens(seed + 10 + i, n, d.vocab) for i, n in enumerate(lens)]
    embs = []
    for s in seqs:
        hs = m(input_ids=torch.from_numpy(s.astype(np.int64))[None]).last_hidden_state[0]
        cls = hs[0]
        embs.append((cls / cls.norm().clamp_min(1e-12)).numpy())
    # one padded batch as HF users would run it (attention_mask), to pin varlen == padded semantics
    L = max(lens)
    ids = torch.zeros(len(seqs), L, dtype=torch.int64)
    mask = torch.zeros(len(seqs), L, dtype=torch.int64)
    for i, s in enumerate(seqs):
        ids[i, :len(s)] = torch.from_numpy(s.astype(np.int64))
        mask[i, :len(s)] = 1
    hs = m(input_ids=ids, attention_mask=mask).last_hidden_state[:, 0]
    padded = (hs / hs.norm(dim=-1, keepdim=True).clamp_min(1e-12)).numpy()
    assert np.abs(padded - np.stack(embs)).max() < 1e-5
    np.savez_compressed(os.path.join(OUT, name + ".npz"), seed=seed, std=std, lens=np.array(lens, np.int32),
                        tokens=np.concatenate(seqs), embeddings=np.stack(embs).astype(np.float32))
    print(name, "embeddings", np.stack(embs).shape)


def gen_llama_fullshape(name, d, seed, n_prompt, n_decode, std=0.02, col_stride=31, rows=(0, 1, 63, 64, 255, 510, 511)):
    """A BASELINE-shape case (full hidden / heads / ffn / 128256 vocab, few layers).  Full logit rows would be 0.5 MB
    each, so the fixture keeps every `col_stride`-th vocabulary column of selected prompt rows and of every decode row,
    plus each row's exact argmax and max value (the quantities parity is judged on)."""
    sd = weights.llama_state_dict(d, seed, std)
    m = hf_llama(d, sd)
    del sd
    prompt = weights.random_tokens(seed + 1, n_prompt, d.vocab)
    ids = torch.from_numpy(prompt.astype(np.int64))[None]
    out = m(ids, use_cache=True)
    pl = out.logits[0]
    rows = [r for r in rows if r < n_prompt]
    cols = np.arange(0, d.vocab, col_stride)
    past = out.past_key_values
    logits = pl[-1]
    toks, srows = [], []
    for _ in range(n_decode):
        srows.append(logits.numpy().copy())
        t = int(torch.argmax(logits))
        toks.append(t)
        o = m(torch.tensor([[t]]), past_key_values=past, use_cache=True)
        past = o.past_key_values
        logits = o.logits[0, -1]
    srows = np.stack(srows)
    np.savez_compressed(os.path.join(OUT, name + ".npz"), seed=seed, std=std, layers=d.layers, prompt=prompt,
                        rows=np.array(rows, np.int32), cols=cols.astype(np.int32),
                        prompt_logits=pl[rows][:, cols].numpy().astype(np.float32),
                        prompt_argmax=pl.argmax(-1).numpy().astype(np.int32), prompt_max=pl.max(-1).values.numpy().astype(np.float32),
                        prompt_absmax=pl.abs().max(-1).values.numpy().astype(np.float32),
                        greedy_tokens=np.array(toks, np.int32), step_logits=srows[:, cols].astype(np.float32),
                        step_max=srows.max(-1).astype(np.float32))
    print(name, "greedy", toks)


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "full":  # the two BASELINE-shape fixtures (minutes, ~10 GB of RAM)
        gen_bert("bge_base_full", configs.bge_base(), 2, [512, 1, 64, 129, 300, 511], 0.02)
        d8 = configs.llama3_8b()
        d8.layers = 2
        gen_llama_fullshape("llama3_8b_2layer", d8, 4, 512, 8)
        sys.exit(0)
    if len(sys.argv) > 1 and sys.argv[1] == "qwen2":
        gen_qwen2("qwen2_tiny", configs.tiny_qwen2(layers=2, vocab=1000), 6, 70, 10)
        sys.exit(0)
    gen_llama("llama_tiny_d64", configs.tiny_llama(layers=2, head_dim=64, vocab=1000), 0, 48, 12, 0.02)
    gen_llama("llama_tiny_d64_s05", configs.tiny_llama(layers=2, head_dim=64, vocab=1000), 0, 48, 12, 0.05)
    gen_llama("llama_tiny_d128_rope3", configs.tiny_llama(layers=3, head_dim=128, vocab=1000, rope_scaling=True), 3, 200, 8, 0.05)
    gen_bert("bert_tiny", configs.tiny_bert(layers=2, vocab=1000), 5, [7, 64, 129, 300, 1, 512])
