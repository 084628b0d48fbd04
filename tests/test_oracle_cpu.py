"""CPU: the oracle restatement against the golden vectors produced by the reference's own code
(oracle/make_golden.py), and its unit properties."""
import numpy as np
import pytest
import torch

from conftest import load_golden, model_and_weights, pcm_to_wave
from oracle import quartznet_oracle as O

CASES = [("vi12x1", "rand"), ("en15x5", "rand"), ("vi12x1", "real_batch"), ("vi12x1", "real_single"),
         ("en15x5", "real_batch")]


@pytest.mark.parametrize("tag,kind", CASES)
def test_oracle_matches_reference_golden(tag, kind):
    torch.set_num_threads(8)
    md, enc_sd, dec_sd = model_and_weights(tag, "rand" if kind == "rand" else "real")
    g = load_golden(f"{tag}_{kind}")
    r = O.full_path(enc_sd, dec_sd, md["JasperEncoder"]["jasper"], pcm_to_wave(g["pcm16"]), torch.from_numpy(g["lens"]))
    np.testing.assert_allclose(r["feats"].numpy(), g["feats"], atol=2e-5, rtol=0)
    assert r["seq"].tolist() == g["seq"].tolist()
    assert r["enc_len"].tolist() == g["enc_len"].tolist()
    np.testing.assert_allclose(r["enc"][:, ::32, :].numpy(), g["enc_sub"], atol=2e-4, rtol=0)
    rel = np.linalg.norm(r["logits"].numpy() - g["logits"]) / np.linalg.norm(g["logits"])
    assert rel < 1e-5, rel
    top2 = torch.from_numpy(g["logits"]).log_softmax(-1).topk(2, -1).values
    safe = (top2[..., 0] - top2[..., 1]) > 1e-3
    assert torch.equal(r["ids"][safe], torch.from_numpy(g["ids"])[safe])
    texts = O.ids_to_text(O.ctc_collapse(g["ids"], len(md["labels"])), md["labels"])
    assert texts == [str(t) for t in g["texts"]]


def test_mel_basis_matches_torchaudio_slaney():
    torchaudio = pytest.importorskip("torchaudio")
    fb = O.slaney_mel_filterbank(16000, 512, 64, 0.0, 8000.0)
    ta = torchaudio.functional.melscale_fbanks(257, 0.0, 8000.0, 64, 16000, norm="slaney", mel_scale="slaney").T.numpy()
    assert np.abs(fb - ta).max() / np.abs(ta).max() < 1e-5
    nnz = np.count_nonzero(fb, axis=1)
    assert nnz.min() >= 2 and nnz.max() <= 23


def test_ctc_collapse_rules():
    blank = 3
    ids = np.array([[3, 3, 3, 3], [0, 0, 1, 1], [0, 3, 0, 0], [2, 2, 2, 2], [3, 1, 3, 1]])
    assert O.ctc_collapse(ids, blank) == [[], [0, 1], [0, 0], [2], [1, 1]]


def test_length_chain_is_fractional_then_truncated():
    # seq=500 -> dw(stride 2) 250.5 -> pw masks at 250 -> 250.0 (parts/jasper.py:108-121)
    x = torch.zeros(1, 4, 501)
    w = torch.zeros(4, 1, 33)
    out, lens = O.masked_conv1d(x, torch.tensor([500]), w, stride=2, padding=16, dilation=1, groups=4)
    assert out.shape[-1] == 251 and lens.item() == 250.5
    _, lens2 = O.masked_conv1d(out, lens, torch.zeros(8, 4, 1))
    assert lens2.item() == 250.0


def test_beam_oracle_properties():
    """oracle/beam_oracle.py (restatement of pyctcdecode without LM - parity unpinned): beam width 1 on a peaked
    posterior equals the greedy collapse; wider beams never score worse; whitespace is normalised."""
    from oracle import beam_oracle as BO
    g = load_golden("vi12x1_real_single")
    import viet_asr_b200 as V
    labels = V.configs.VI_LABELS
    lp = torch.from_numpy(g["logits"]).log_softmax(-1)[0].numpy()
    t1, s1 = BO.beam_search_no_lm(lp, labels, 1)
    t20, s20 = BO.beam_search_no_lm(lp, labels, 20)
    greedy = O.ids_to_text(O.ctc_collapse(g["ids"], len(labels)), labels)[0]
    assert t1 == " ".join(greedy.split())
    assert s20 >= s1 - 1e-9 and t20 == t1
    en = V.configs.EN_LABELS
    T, V1 = 12, len(en) + 1
    x = np.full((T, V1), -20.0, dtype=np.float32)
    for t, c in enumerate([0, 0, 8, 28, 9, 0, 28, 0, 20, 28, 28, 28]):      # "  hi  t" with blanks
        x[t, c] = 0.0
    x = torch.from_numpy(x).log_softmax(-1).numpy()
    assert BO.beam_search_no_lm(x, en, 8)[0] == "hi t"


# ----------------------------------------------------------------------------- KenLM binary (oracle + product parser)
@pytest.mark.parametrize("order", [3, 5])
def test_kenlm_readers_agree_on_tiny_fixture(order):
    """oracle/kenlm_oracle.py (scalar bit reads) and viet-asr_b200/kenlm_binary.py (vectorised decode -> flat arrays
    for the GPU) are independent restatements of the KenLM QUANT_ARRAY_TRIE layout; they must agree on every node."""
    from conftest import tiny_lm_path
    from oracle.kenlm_oracle import KenlmBinary
    from viet_asr_b200.kenlm_binary import KenlmModel
    a, m = KenlmBinary(tiny_lm_path(order)), KenlmModel(tiny_lm_path(order))
    assert a.order == m.order == order and a.counts == m.counts and a.words == m.words
    assert (a.bos, a.eos) == (m.bos, m.eos) and a.words[0] == "<unk>"
    for w in range(a.counts[0]):
        p, b, lo, hi = a.unigram(w)
        assert (m.uni_prob[w], m.uni_backoff[w], m.uni_next[w], m.uni_next[w + 1]) == (np.float32(p), np.float32(b), lo, hi)
    for k in range(order - 2):
        for i in range(a.counts[k + 1]):
            p, b, lo, hi = a.middle(k, i)
            assert m.mid_word[k][i] == a.middle_word(k, i)
            assert (m.mid_prob[k][i], m.mid_backoff[k][i], m.mid_next[k][i], m.mid_next[k][i + 1]) == (np.float32(p), np.float32(b), lo, hi)
    for i in range(a.counts[-1]):
        assert m.long_word[i] == a.longest_word(i) and m.long_prob[i] == np.float32(a.longest(i))


def test_kenlm_oracle_backoff_semantics_on_tiny_fixture():
    """score() = prob of the longest matching n-gram + back-offs of the unmatched context suffixes."""
    from conftest import tiny_lm_path
    from oracle.kenlm_oracle import KenlmBinary
    from oracle.kenlm_writer import tiny_ngrams
    lm = KenlmBinary(tiny_lm_path(3))
    grams = tiny_ngrams(3, 11)
    tri = [g for g in grams if len(g) == 3][:50]
    for g in tri:                                              # a stored trigram scores as its own (quantised) prob
        ids = [lm.index(w) for w in g]
        found = lm.walk(ids[::-1])
        assert len(found) == 3 and lm.score(ids[:2], ids[2]) == found[-1][0]
        assert abs(found[-1][0] - grams[g][0]) < 0.15          # snapped to the nearest of 256 random bins over [-6, 0]
    # unseen word after a seen bigram context: bigram/unigram prob + the context's back-offs
    ctx = [lm.index("the"), lm.index("cat")]
    w = 0                                                       # <unk>: no bigram or trigram ends in it
    uni = lm.unigram(w)
    cw = lm.walk(ctx[::-1])
    assert lm.score(ctx, w) == uni[0] + sum(b for _, b in cw)
    assert lm.index("zebra") == 0 and lm.score_words(["zebra"]) == lm.score([lm.bos], 0) + lm.score([lm.bos, 0], lm.eos)


@pytest.mark.parametrize("name", ["3-gram-lm.binary", "5-gram-lm.binary"])
def test_kenlm_shipped_model_is_understood(name):
    """What pins the layout restatement: the reference's own files.  Section sizes add up (checked by the readers),
    the stored vocabulary hashes are MurmurHash64A of the strings in id order, child ranges are monotone (checked by
    the product parser) and P(. | context) sums to 1 (to quantisation noise) for contexts of every length."""
    from conftest import shipped_lm_path
    from oracle.kenlm_oracle import KenlmBinary
    from oracle.kenlm_writer import murmur64a
    from viet_asr_b200.kenlm_binary import KenlmModel
    path = shipped_lm_path(name)
    lm = KenlmBinary(path)
    KenlmModel(path)                                            # raises if its structural checks fail
    off = (108 + 8 * lm.order + 7) // 8 * 8 + 8
    h = np.frombuffer(lm.d, dtype="<u8", count=lm.counts[0] - 1, offset=off)
    assert all(murmur64a(lm.words[i + 1].encode("utf-8")) == int(h[i]) for i in range(0, len(h), 7))
    rng = np.random.default_rng(0)
    words = [w for w in ("không", "cần", "phải", "và", "là", "của") if w in lm.word2id]
    ctxs = [[], [lm.bos]] + [[lm.index(w) for w in words[i:i + n]] for n in (1, 2, 3, 4) for i in (0, 1)]
    for ctx in ctxs:
        total = sum(10.0 ** lm.score(ctx, w) for w in range(lm.counts[0]) if w != lm.bos)
        assert abs(total - 1.0) < 0.03, (ctx, total)
    assert lm.score_words("không cần phải".split()) > lm.score_words("phải không cần cần".split())


def test_beam_lm_oracle_properties():
    """oracle/beam_oracle.beam_search_lm (restatement of pyctcdecode with a KenLM model - parity unpinned):
    alpha = beta = unk = 0 reduces to the search without LM; the LM pulls an acoustically ambiguous word towards the
    vocabulary; the combined score is acoustic + LM."""
    from conftest import spelled_posteriors, tiny_lm_path
    from oracle import beam_oracle as BO
    from oracle.kenlm_oracle import KenlmBinary
    import viet_asr_b200 as V
    labels = V.configs.EN_LABELS
    lm = KenlmBinary(tiny_lm_path(3))
    lp = spelled_posteriors(["the cat sat on the mat", "hi there"], labels, seed=5, noise=1.0).numpy()
    for x in lp:
        t0, s0 = BO.beam_search_no_lm(x, labels, 16)
        t1, s1 = BO.beam_search_lm(x, labels, 16, lm, alpha=0.0, beta=0.0, unk_score_offset=0.0)
        assert t0 == t1 and abs(s0 - s1) < 1e-9
    # 'cat' vs 'cxt' nearly tied acoustically: only the LM (unknown-word penalty) separates them
    amb = spelled_posteriors(["the cat sat"], labels, seed=9, noise=0.2, confusions=[(0, 5, "x")]).numpy()[0]
    with_lm = BO.beam_search_lm(amb, labels, 16, lm)[0]
    assert with_lm == "the cat sat"
    allb = BO.beam_search_lm(amb, labels, 16, lm, return_all=True)
    assert allb[0][2] == max(b[2] for b in allb)


def test_conv_stft_restatement_equals_fft_with_periodic_window():
    """`stft_conv: true` (features.py:156-167): the restated torch_stft convolution STFT (parity with the package
    unpinned) must equal an FFT-based STFT with scipy's periodic window - the identity the CUDA front end relies on."""
    g = torch.Generator().manual_seed(3)
    x = 0.1 * torch.randn(2, 4000, generator=g)
    for name, fn in (("hann", torch.hann_window), ("hamming", torch.hamming_window)):
        mag = O.conv_stft_magnitude(x, 512, 160, 320, name)
        ref = torch.stft(x, 512, 160, 320, window=fn(320, periodic=True), center=True, return_complex=True).abs()
        assert mag.shape == ref.shape == (2, 257, 26)
        assert (mag - ref).abs().max().item() < 2e-4 * ref.abs().max().item()
    a, _ = O.filterbank_features(x, torch.tensor([4000, 3000]), stft_conv=True)
    b, _ = O.filterbank_features(x, torch.tensor([4000, 3000]), stft_conv=False)
    assert a.shape == b.shape and 1e-4 < (a - b).abs().max().item() < 0.5      # periodic vs symmetric window: close, not equal


def test_c_restatement_of_argmax_and_collapse():
    """oracle/c/ctc_oracle.c (plain C) == the numpy/torch restatement == the reference's texts on the golden vectors,
    plus ties (first maximal index) and the all-frames rule."""
    from oracle import c_oracle as CO
    for tag, kind in CASES:
        g = load_golden(f"{tag}_{kind}")
        logp = torch.from_numpy(g["logits"]).log_softmax(-1).numpy()
        ids = CO.greedy_argmax(logp)
        assert np.array_equal(ids, torch.from_numpy(logp).argmax(-1).numpy())
        blank = logp.shape[-1] - 1
        out, n = CO.ctc_collapse(g["ids"], blank)
        want = O.ctc_collapse(g["ids"], blank)
        assert [row[:k].tolist() for row, k in zip(out, n)] == want
        assert all((row[k:] == -1).all() for row, k in zip(out, n))
        md = model_and_weights(tag, "rand")[0]
        assert O.ids_to_text([row[:k].tolist() for row, k in zip(out, n)], md["labels"]) == [str(t) for t in g["texts"]]
    tie = np.zeros((1, 3, 5), np.float32); tie[0, 1, 2] = tie[0, 1, 4] = 1.0; tie[0, 2, 3] = np.nan
    assert CO.greedy_argmax(tie).tolist() == [[0, 2, 3]] == torch.from_numpy(tie).argmax(-1).tolist()
    out, n = CO.ctc_collapse(np.array([[3, 3, 3, 3], [0, 0, 1, 1], [0, 3, 0, 0], [3, 1, 3, 1]]), 3)
    assert n.tolist() == [0, 2, 2, 2] and out[1, :2].tolist() == [0, 1] and out[2, :2].tolist() == [0, 0]


def test_oracle_matches_round2_goldens():
    """oracle/make_golden_r2.py fixtures (outputs of the reference's own parts/jasper.py): the restatement reproduces
    the reference ids on a sample transcribed alone (vi 12x1) and on two clips of the benchmark-shaped 15x5 batch."""
    from conftest import have_weights, load_weights
    import viet_asr_b200 as V
    if not have_weights("vi12x1") or not have_weights("en15x5"):
        pytest.skip("shipped checkpoints not present")
    g = load_golden("vi12x1_real_all")
    md = V.configs.quartznet12x1_vi()
    enc_sd, dec_sd = load_weights("vi12x1")
    i = 5                                                   # the shortest sample (2.4 s)
    n, f = int(g["lens"][i]), int(g["frames"][i])
    r = O.full_path(enc_sd, dec_sd, md["JasperEncoder"]["jasper"], pcm_to_wave(g["pcm16"][i:i + 1, :n]), torch.tensor([n]))
    assert r["ids"].shape[1] == f == O.utterance_frames(md["JasperEncoder"]["jasper"], [n])[0]
    assert np.array_equal(r["ids"][0].numpy(), g["ids"][i, :f])
    assert O.ids_to_text(O.ctc_collapse(r["ids"].numpy(), len(md["labels"])), md["labels"]) == [str(g["texts"][i])]
    b = load_golden("en15x5_real_b48")
    md = V.configs.quartznet15x5()
    enc_sd, dec_sd = load_weights("en15x5")
    L = int(b["L"])
    clips = []
    for s_, o in list(zip(b["src"], b["off"]))[:2]:
        x = np.roll(g["pcm16"][s_, : int(g["lens"][s_])], -int(o))
        clips.append(np.tile(x, -(-L // len(x)))[:L])
    r = O.full_path(enc_sd, dec_sd, md["JasperEncoder"]["jasper"], pcm_to_wave(np.stack(clips)), torch.full((2,), L))
    assert np.array_equal(r["ids"].numpy(), b["ids"][:2].astype(np.int64))
    ref_logp = torch.from_numpy(b["logits4"][:2]).log_softmax(-1)
    assert (r["logp"] - ref_logp).abs().max().item() < 2e-4


def test_feature_noise_floor_of_the_reference_itself():
    """How far the reference's own fp32 features are from a float64 evaluation of the same formulas
    (features.py:245-301): 5e-5 on the real-speech fixture, up to 9e-5 on white noise.  This is the floor any fp32
    implementation with a different FFT order sits on; the GPU tests bound |kernel - reference| by FEAT_ATOL = 2e-3
    (about 20x this floor; the binding checks downstream are log-probs <= 1e-3 rel-L2 and bit-exact greedy ids)."""
    g = load_golden("vi12x1_real_batch")
    wave, length = pcm_to_wave(g["pcm16"]), torch.from_numpy(g["lens"])
    ref, seq = O.filterbank_features(wave, length)
    x = wave.double()
    x = torch.cat((x[:, 0].unsqueeze(1), x[:, 1:] - 0.97 * x[:, :-1]), dim=1)
    win = torch.hann_window(320, periodic=False, dtype=torch.float64)
    spec = torch.stft(x, n_fft=512, hop_length=160, win_length=320, center=True, window=win, return_complex=True)
    power = torch.view_as_real(spec).pow(2).sum(-1)
    fb = torch.from_numpy(O.slaney_mel_filterbank(16000, 512, 64, 0.0, 8000.0)).double().unsqueeze(0)
    logmel = torch.log(torch.matmul(fb, power) + 2.0 ** -24)
    exact = torch.zeros_like(logmel)
    for b in range(x.shape[0]):
        n = int(seq[b])
        seg = logmel[b, :, :n]
        exact[b, :, :n] = (seg - seg.mean(1, keepdim=True)) / (seg.std(1, keepdim=True) + 1e-5)
    floor = (ref.double() - exact).abs().max().item()
    assert 1e-6 < floor < 2e-4, floor
