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
