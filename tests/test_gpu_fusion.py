"""GPU parity of the label-fusion path (platipy/imaging/label/fusion.py) against the CPU oracle."""
import numpy as np
import pytest

from oracle import itk_oracle as orc
from oracle import platipy_ref as ref
from platipy_b200 import fusion
from platipy_b200 import sitk_compat as sk
from platipy_b200.sitk_compat import Image
from platipy_b200.synth import synth_labels, synth_pair

pytestmark = pytest.mark.gpu
REL_TOL = 1e-5


def _atlas_set(n_atlas, size, n_struct, weights, missing=None):
    rng = np.random.default_rng(5)
    sp = (1.0, 1.0, 1.5)
    atlas_set = {}
    base = synth_labels(size, n_struct, seed=400)
    for a in range(n_atlas):
        entry = {}
        for s in range(n_struct):
            if missing and (a, s) in missing:
                continue
            lab = np.roll(base[s], shift=(a % 3 - 1, (a * 2) % 5 - 2, a % 4 - 1), axis=(0, 1, 2))
            entry[f"S{s}"] = Image(lab.astype(np.uint8), sp)
        entry["Weight Map"] = Image(weights(a, size, rng), sp)
        atlas_set[f"{a:03d}"] = {"DIR": entry}
    return atlas_set


def test_weight_maps_match_oracle(engine):
    t, m = synth_pair((40, 36, 20), seed=51, spacing=(1.0, 1.0, 2.0))
    params = {"sigma": 2.0, "epsilon": 1e-5, "factor": 1e12, "normalise": False}
    for vt in ("unweighted", "global", "local"):
        got = fusion.compute_weight_map(t, m, vt, params)
        exp = ref.compute_weight_map(t, m, vt, params)
        assert got.GetPixelID() == sk.sitkFloat32
        if vt == "global":  # one scalar from a 28k-term double sum: reduction order differs
            assert np.allclose(got.array, exp.array, rtol=1e-6, atol=0)
        elif vt == "local":  # pow(x, -1) on the device is not bit-identical to libm
            assert np.allclose(got.array, exp.array, rtol=REL_TOL, atol=0)
        else:
            assert np.array_equal(got.array, exp.array)
    assert np.array_equal(fusion.compute_weight_map(t, m, "unweighted", None).array, np.ones(t.array.shape, np.float32))
    with pytest.raises(UnboundLocalError):  # the reference falls through its if / elif chain (fusion.py:151-202)
        fusion.compute_weight_map(t, m, "no_such_vote", params)


def test_block_weight_map_and_normalise_match_oracle(engine):
    """vote_type "block" (fusion.py:179-200: BoxMean -> Pow -> ** |gain/2| -> * factor) and the normalise option
    (bool / mask image, fusion.py:171-177,196-200).  Float32 stages; tolerance = the north star's 1e-5 relative
    (device pow() is not bit-identical to libm, the box sums differ from ITK's integral image by double rounding)."""
    t, m = synth_pair((40, 36, 20), seed=52, spacing=(1.0, 1.0, 2.0))
    mask = Image((np.random.default_rng(2).random(t.array.shape) > 0.7).astype(np.uint8), t.GetSpacing())
    for block, gain, normalise in ((5, 6, False), ((2, 3, 1), 4, False), (3, 6, True), (1, 2, mask), (0, 6, False)):
        params = {"factor": 1e12, "gain": gain, "blockSize": block, "normalise": normalise}
        got = fusion.compute_weight_map(t, m, "block", params)
        exp = ref.compute_weight_map(t, m, "block", params)
        assert got.GetPixelID() == sk.sitkFloat32
        assert np.all(np.isfinite(exp.array))
        assert np.allclose(got.array, exp.array, rtol=REL_TOL, atol=0), (block, gain)
    for normalise in (True, mask):
        params = {"sigma": 2.0, "epsilon": 1e-5, "normalise": normalise}
        got = fusion.compute_weight_map(t, m, "local", params)
        exp = ref.compute_weight_map(t, m, "local", params)
        assert np.allclose(got.array, exp.array, rtol=REL_TOL, atol=0)
        if normalise is True:
            assert got.array.max() == 1.0


def test_combine_labels_bit_exact(engine):
    size = (36, 30, 22)
    flat = _atlas_set(5, size, 3, lambda a, s, rng: np.ones(s[::-1], np.float32))
    rnd = _atlas_set(4, size, 2, lambda a, s, rng: (rng.random(s[::-1]) * (rng.random(s[::-1]) > 0.2)).astype(np.float32), missing={(1, 1)})
    for atlas_set, names in ((flat, ["S0", "S1", "S2"]), (rnd, ["S0", "S1"]), (flat, "S1")):
        got = fusion.combine_labels(atlas_set, names)
        exp = ref.combine_labels(atlas_set, names)
        assert list(got) == list(exp)
        for k in got:
            assert got[k].GetPixelID() == sk.sitkFloat32
            assert np.array_equal(got[k].array, exp[k].array), k
            assert got[k].array.max() == 1.0 and np.all((got[k].array == 0) | (got[k].array >= np.float32(1e-4)))
    # threshold=0 / different smoothing
    got = fusion.combine_labels(flat, "S0", threshold=0, smooth_sigma=2.0)
    exp = ref.combine_labels(flat, "S0", threshold=0, smooth_sigma=2.0)
    assert np.array_equal(got["S0"].array, exp["S0"].array)


def test_staple_matches_oracle(engine):
    size = (40, 32, 24)
    truth = synth_labels(size, 1, seed=77)[0]
    rng = np.random.default_rng(6)
    raters = {}
    for k in range(7):
        r = truth.copy()
        r[rng.random(truth.shape) < 0.02 * (k + 1)] ^= 1
        raters[str(k)] = {"HEART": Image(r.astype(np.float32) * (1.0 if k % 2 else 0.75)), "OTHER": Image(np.roll(r, k, axis=2))}
    got = fusion.combine_labels_staple(raters)
    exp = ref.combine_labels_staple(raters)
    assert sorted(got) == sorted(exp) == ["HEART", "OTHER"]
    for k in got:
        assert got[k].GetPixelID() == sk.sitkFloat64
        # serial double sums over the whole volume on the CPU vs tree sums on the GPU: p, q agree to ~1e-15
        assert np.allclose(got[k].array, exp[k].array, rtol=1e-9, atol=1e-12), k
        assert got[k].array.max() == 1.0 and np.all((got[k].array == 0) | (got[k].array >= 1e-4))
    # p / q / iteration count of the EM itself
    dec = [(raters[k]["HEART"].array >= 0.5).astype(np.uint8) for k in raters]
    W, p, q, it = orc.staple(dec)
    gW, info = engine.staple([engine.to_device(Image(d)) for d in dec], threshold=0.0, rescale=False)
    assert info["elapsed_iterations"] == it
    assert np.allclose(info["p"], p, rtol=1e-12) and np.allclose(info["q"], q, rtol=1e-12)
    assert np.allclose(engine.to_host(gW, pinned=False).array, W, rtol=1e-10, atol=1e-14)


def test_compact_exchange_formats_match_the_plain_paths(engine):
    """The structure-sharded fusion's payloads (platipy_b200/multiatlas.py): STAPLE straight from the packed decision mask equals
    STAPLE on the decision volumes (same histogram, same table EM: bit-identical) -- dense and sparse holder masks, 8- and 16-bit
    masks -- and equals the oracle to the reduction-order tolerance; the UInt8 vote counts give the float32 votes bit for bit."""
    import torch

    eng = engine
    size, sp = (44, 36, 20), (1.0, 1.0, 1.5)
    base = synth_labels(size, 1, seed=400)[0]
    n_atlas = 11
    labs = [eng.to_device(Image(np.roll(base, shift=(a % 3 - 1, (a * 2) % 5 - 2, a % 4 - 1), axis=(0, 1, 2)).astype(np.uint8), sp)) for a in range(n_atlas)]
    for holders, dtype in (([0, 1, 2, 3, 4], torch.uint8), ([0, 2, 3, 7], torch.uint8), ([1, 4, 8, 9, 10], torch.int16), (list(range(11)), torch.int32)):
        with torch.cuda.stream(eng.stream):
            packed = torch.zeros(labs[0].tensor.shape, dtype=dtype, device=eng.device)
        for a in holders:
            eng.pack_label(labs[a], a, packed, False)
        hm = sum(1 << a for a in holders)
        got, info = eng.staple_packed(packed, hm, labs[0], threshold=1e-4, rescale=True, want_info=True)
        plain, info_p = eng.staple([labs[a] for a in holders], threshold=1e-4, rescale=True)
        g, p = eng.to_host(got).array, eng.to_host(plain).array
        assert np.array_equal(g, p) and info["elapsed_iterations"] == info_p["elapsed_iterations"] and info["p"] == info_p["p"]
        W, po, qo, it = orc.staple([eng.to_host(labs[a]).array for a in holders])
        exp = orc.rescale_threshold_f64(W, 1e-4)
        assert it == info["elapsed_iterations"]
        assert np.allclose(g, exp, rtol=1e-9, atol=1e-12)
        # without the statistics the call does not synchronise; same values
        assert np.array_equal(eng.to_host(eng.staple_packed(packed, hm, labs[0])).array, g)
    # unweighted vote: counts vs the float32 accumulators
    with torch.cuda.stream(eng.stream):
        counts = torch.zeros(labs[0].tensor.shape, dtype=torch.uint8, device=eng.device)
        flag = torch.zeros(1, dtype=torch.int32, device=eng.device)
    num, den = eng.empty(labs[0].tensor.shape, np.float32), eng.empty(labs[0].tensor.shape, np.float32)
    ones = eng.weight_map(eng.cast(labs[0], np.float32), eng.cast(labs[0], np.float32), 0)
    for k, a in enumerate(range(7)):
        eng.count_accumulate(labs[a], counts, False, flag)
        eng.vote_accumulate(labs[a], ones, num, den, k == 0)
    got = eng.to_host(eng.vote_finalize_counts(counts, 7, labs[0], 1.0, 1e-4)).array
    exp = eng.to_host(eng.vote_finalize(num, den, labs[0], 1.0, 1e-4)).array
    assert int(flag.item()) == 0 and np.array_equal(got, exp)
    assert np.array_equal(got, orc.combine_labels_f32([eng.to_host(labs[a]).array for a in range(7)], [np.ones(base.shape, np.float32)] * 7,
                                                      orc.geom_of(labs[0]), 1.0, 1e-4))
    # a label value above 1 raises the flag (the caller then takes the float32 path)
    two = eng.to_device(Image((base * 2).astype(np.uint8), sp))
    eng.count_accumulate(two, counts, False, flag)
    assert int(flag.item()) == 1
    with pytest.raises(ValueError):
        eng.pack_label(labs[0], 9, torch.zeros(labs[0].tensor.shape, dtype=torch.uint8, device=eng.device), False)
