"""CPU: the native tokenizer (hb_tok_*, helix_b200/csrc/tokenizer.cpp) against HF `tokenizers` 0.22 — bit-exact token ids.
No vocabulary files exist offline, so the test TRAINS byte-level BPE vocabularies with the HF library on a synthetic
multilingual corpus using the Llama-3 pipeline (Split(regex, isolated) + ByteLevel(use_regex=False), `ignore_merges` on and
off, Llama-3's special tokens), saves tokenizer.json, loads it natively and compares ids, decoded text and the chat template
on thousands of generated strings (hypothesis) plus hand-picked pre-tokenizer edge cases."""
import json
import random

import pytest
from hypothesis import given, settings, strategies as st
from tokenizers import Regex, Tokenizer, decoders, models, normalizers, pre_tokenizers, processors, trainers

from helix_b200.server import HFTokenizer
from helix_b200.tokenizer import NativeTokenizer

LLAMA3_SPLIT = (r"(?i:'s|'t|'re|'ve|'m|'ll|'d)|[^\r\n\p{L}\p{N}]?\p{L}+|\p{N}{1,3}| ?[^\s\p{L}\p{N}]+[\r\n]*|\s*[\r\n]+|\s+(?!\S)|\s+")
SPECIALS = ["<|begin_of_text|>", "<|end_of_text|>", "<|start_header_id|>", "<|end_header_id|>", "<|eot_id|>"]

WORDS = ("the of and to in is that for it as with was on be by at this have from or an they which one you were her all she "
         "there would their we him been has when who will more no if out so said what up its about into than them can only "
         "other new some could time these two may then do first any my now such like our over man me even most made after also "
         "did many before must through back years where much your way well down should because each just those people mr how "
         "too little state good very make world still own see men work long get here between both life being under never day "
         "naïve café über straße señor façade coöperate Ελληνικά привет мир здравствуйте 日本語 中文字符 한국어 שלום مرحبا "
         "I'm you're they've we'll he'd it's don't DON'T I'LL 3.14159 2024 100000 7 42 0x1F a_b snake_case CamelCase "
         "e-mail https://example.com/path?q=1&r=2 foo@bar.org #tag @user $9.99 50% (parens) [brackets] {braces} <tags> "
         "∑ ∞ ≠ → ← ✓ ✗ 😀 🎉 👍🏽 ♥ © ® ™ … — – “quoted” ‘single’ «guillemets»").split()


def corpus(seed, n=4000):
    rnd = random.Random(seed)
    seps = [" ", " ", " ", "  ", "\n", "\n\n", "\t", " \n", ", ", ". ", "! ", "? ", ": ", "; ", " - ", "\r\n", "   ", "\n  ", ""]
    out = []
    for _ in range(n):
        k = rnd.randint(1, 24)
        out.append("".join(rnd.choice(WORDS) + rnd.choice(seps) for _ in range(k)))
    return out


def train(tmp_path, ignore_merges, vocab_size, seed):
    tok = Tokenizer(models.BPE(ignore_merges=ignore_merges))
    tok.pre_tokenizer = pre_tokenizers.Sequence([pre_tokenizers.Split(Regex(LLAMA3_SPLIT), behavior="isolated", invert=False),
                                                 pre_tokenizers.ByteLevel(add_prefix_space=False, trim_offsets=True, use_regex=False)])
    tok.decoder = decoders.ByteLevel()
    trainer = trainers.BpeTrainer(vocab_size=vocab_size, special_tokens=SPECIALS, initial_alphabet=pre_tokenizers.ByteLevel.alphabet(),
                                  show_progress=False)
    tok.train_from_iterator(corpus(seed), trainer)
    path = tmp_path / f"tokenizer_{int(ignore_merges)}_{vocab_size}.json"
    tok.save(str(path))
    if ignore_merges:  # make sure the flag reached the file (older trainers drop it)
        d = json.loads(path.read_text())
        d["model"]["ignore_merges"] = True
        path.write_text(json.dumps(d))
        tok = Tokenizer.from_file(str(path))
    return tok, path


EDGE = ["", " ", "  ", "   x", "x   ", "\n", "\n\n\n", " \n \n x", "a\r\nb", "it's IT'S i'Ll 'tis ''s 'q", "123456789 12 1 1a2b 1,234.50", "¹²³ ½ ٣٤٥ 一二三",
        "hello   world  \n  next", "tab\there", "a b c　d", "!!!\n\n???", " !!! x", "…—“”", "日本語のテキスト。中文，标点！", "mixed日本1st",
        "привет,мир!Ελλάδα", "😀😀 👍🏽\n", "x" * 300, " " * 50 + "y", "\n" * 40, "a'b 'c' d'", "end with space ", "end with newline\n", "\tlead", "1 2  3   4",
        "<|eot_id|> literal specials <|begin_of_text|> inside", "under_score __init__ a__b", "λx.x→y", "'S 'T 'RE 'Ve 'M 'lL 'D 'x"]


@pytest.mark.parametrize("ignore_merges,vocab_size,seed", [(False, 600, 1), (True, 2000, 2)])
def test_ids_and_text_match_hf_tokenizers(tmp_path, ignore_merges, vocab_size, seed):
    hf, path = train(tmp_path, ignore_merges, vocab_size, seed)
    nt = NativeTokenizer(path)
    assert nt.vocab_size == hf.get_vocab_size() and nt.EOS == hf.token_to_id("<|eot_id|>")

    def same(text):
        want = hf.encode(text, add_special_tokens=False).ids
        got = nt.encode(text, parse_special=True)           # HF matches special-token text inside the input too
        assert got == want, (text, got[:20], want[:20])
        assert nt.decode(got, skip_special=False) == hf.decode(want, skip_special_tokens=False)
        assert nt.decode(got, skip_special=True) == hf.decode(want, skip_special_tokens=True)
    for t in EDGE + corpus(seed + 100, 300):
        same(t)
    plain = "say <|eot_id|> please"
    assert nt.encode(plain, parse_special=False) != nt.encode(plain, parse_special=True)   # user text can opt out of specials
    assert nt.decode(nt.encode(plain, parse_special=False)) == plain

    alphabet = st.sampled_from(list("abcXYZ 019'\n\t\r-_.,!?") + ["é", "ß", "Ω", "я", "日", "本", "한", "ע", "م", "😀", "👍", "🏽", " ", " ", "…", "—", "“",
                                                                 "'s", "'LL", "  ", "\n\n", "123", " the", "ing"])

    @settings(max_examples=1500, deadline=None)
    @given(st.lists(alphabet, max_size=40))
    def prop(parts):
        same("".join(parts))
    prop()


def test_llama3_chat_template_matches_the_python_mirror(tmp_path):
    hf, path = train(tmp_path, True, 1500, 3)
    nt, py = NativeTokenizer(path), HFTokenizer(str(path))
    convo = [{"role": "system", "content": "You are terse.\n"}, {"role": "user", "content": "  What's 2+2?\n\nAnswer:"},
             {"role": "assistant", "content": "4"}, {"role": "user", "content": "\nand 3×3? 日本語で"}]
    for msgs in (convo[:1], convo[:2], convo, [{"role": "user", "content": ""}]):
        assert nt.chat(msgs) == py.chat(msgs)
    ids = nt.chat(convo)
    assert ids[0] == nt.token_id("<|begin_of_text|>") and ids.count(nt.EOS) == len(convo)
    assert nt.decode(ids, skip_special=True) == py.decode(ids)


def train_wordpiece(tmp_path, lowercase, seed):
    tok = Tokenizer(models.WordPiece(unk_token="[UNK]", max_input_chars_per_word=100))
    tok.normalizer = normalizers.BertNormalizer(clean_text=True, handle_chinese_chars=True, strip_accents=None, lowercase=lowercase)
    tok.pre_tokenizer = pre_tokenizers.BertPreTokenizer()
    tok.decoder = decoders.WordPiece(prefix="##", cleanup=False)
    trainer = trainers.WordPieceTrainer(vocab_size=900, special_tokens=["[PAD]", "[UNK]", "[CLS]", "[SEP]", "[MASK]"], show_progress=False)
    tok.train_from_iterator(corpus(seed), trainer)
    tok.post_processor = processors.TemplateProcessing(single="[CLS] $A [SEP]", pair="[CLS] $A [SEP] $B:1 [SEP]:1",
                                                       special_tokens=[("[CLS]", tok.token_to_id("[CLS]")), ("[SEP]", tok.token_to_id("[SEP]"))])
    path = tmp_path / f"wordpiece_{int(lowercase)}.json"
    tok.save(str(path))
    return tok, path


@pytest.mark.parametrize("lowercase", [True, False])
def test_wordpiece_ids_match_hf_tokenizers(tmp_path, lowercase):
    """BERT-family encoders (bge): BertNormalizer (clean text, isolate CJK, NFD + strip Mn marks, per-character lowercase)
    -> BertPreTokenizer (whitespace, punctuation isolated) -> WordPiece greedy longest match — ids bit-exact vs HF, with and
    without the [CLS] ... [SEP] framing the embedding path uses."""
    hf, path = train_wordpiece(tmp_path, lowercase, 7)
    nt = NativeTokenizer(path, eos_token="[SEP]")

    def same(text):
        want = hf.encode(text, add_special_tokens=False).ids
        got = nt.encode(text, parse_special=True)
        assert got == want, (text, got[:24], want[:24])
        assert nt.encode_for_embedding(text) == hf.encode(text, add_special_tokens=True).ids
        assert nt.decode(got, skip_special=True) == hf.decode(want, skip_special_tokens=True)
    edge = ["", "  ", "Hello, World!", "naïve café Über STRASSE İstanbul", "日本語のテキスト mixed中文words", "don't stop-believing... 3.14 $9.99 (x) [y] {z}",
            "x" * 101 + " ok", "zzzzqqqqjjjj unknownword", "tab\tsep\nnew\r\nline", "ＦＵＬＬｗｉｄｔｈ １２３", "Ångström Å Å ﬁ ǅ", "한국어 텍스트 가각",
            "[CLS] literal [SEP] [MASK] inside", "emoji 😀 👍🏽 ok", "«quoted» “double” ‘single’ – dash — em", "a\u200bb\u00adc\ufeffd"]
    edge = [e.encode().decode("unicode_escape") if "\\u" in e or "\\t" in e else e for e in edge]
    for t in edge + corpus(107, 300):
        same(t)
    alphabet = st.sampled_from(list("abcXYZ 019'\n\t-_.,!?()") + ["é", "É", "ß", "Ω", "Σ", "я", "Я", "日", "本", "한", "ע", "😀", "\u00a0", "\u2003", "…", "—",
                                                                   "the", "ing", "##", "İ", "ǅ", "ﬁ", "Å", "ü", "Ü", "\u0301", "\u200b"])

    @settings(max_examples=1200, deadline=None)
    @given(st.lists(alphabet, max_size=40))
    def prop(parts):
        same("".join(parts))
    prop()
