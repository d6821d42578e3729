"""BASELINE.json full sizes (512x512x256): what the oracle can check in seconds is checked against it (one
full-resolution Demons iteration, label / image resampling); the rest through size-independent properties
(identical images -> zero field after one iteration; zero displacement -> identity resampling; batched ==
per-image resampling)."""
import numpy as np
import pytest

from oracle import itk_oracle as orc
from platipy_b200 import registration as reg
from platipy_b200 import sitk_compat as sk
from platipy_b200.sitk_compat import Image
from platipy_b200.synth import smooth_random_dvf, synth_labels, synth_pair

pytestmark = pytest.mark.gpu
SIZE = (512, 512, 256)


@pytest.fixture(scope="module")
def pair():
    return synth_pair(SIZE, seed=0, moving_seed=100)


def _params(iters):
    f = reg.FastSymmetricForcesDemonsRegistrationFilter()
    f.SetStandardDeviations((1.5, 1.5, 1.5))
    f.SetSmoothUpdateField(True)
    return f.params(iters)


def test_one_full_resolution_iteration_matches_oracle(engine, pair):
    fixed, moving = pair
    D, st = orc.demons_execute(fixed.array, orc.geom_of(fixed), moving.array, orc.geom_of(moving), orc.demons_params((1.5,) * 3, 2, smooth_update_field=True))
    gD, gst = engine.demons_execute(engine.to_device(fixed), engine.to_device(moving), _params(2))
    got = engine.to_host(gD, pinned=False).array
    assert gst["elapsed_iterations"] == st["elapsed_iterations"] == 2
    assert np.abs(got - D).max() <= 1e-4
    assert np.array_equal(got, D)
    assert abs(gst["metric"] - st["metric"]) <= 1e-9 * st["metric"]


def test_identical_images_give_zero_field(engine, pair):
    fixed, _ = pair
    dF = engine.to_device(fixed)
    img, tfm, dvf = reg.fast_symmetric_forces_demons_registration(dF, dF, resolution_staging=[4, 2, 1], iteration_staging=[100, 50, 25])
    assert [s["elapsed_iterations"] for s in reg.LAST_LEVEL_STATS] == [1, 1, 1]
    assert float(dvf.tensor.abs().max()) == 0.0
    assert bool((img.tensor == dF.tensor).all())


def test_config3_apply_transform_labels_and_image(engine, pair):
    """BASELINE.json configs[2]: CT (linear, -1000) + 20 UInt8 masks (nearest neighbour, 0) through one dense DVF."""
    fixed, moving = pair
    labels = [Image(l) for l in synth_labels(SIZE, 20, seed=200)]
    dvf = Image(smooth_random_dvf(SIZE, seed=9, peak_mm=6.0), is_vector=True)
    tfm = sk.DisplacementFieldTransform(dvf)
    d_imgs = [engine.to_device(moving)] + [engine.to_device(l) for l in labels]
    outs = reg.apply_transform_batch(d_imgs, d_imgs[0], tfm, [-1000] + [0] * 20, [sk.sitkLinear] + [sk.sitkNearestNeighbor] * 20)
    # per-call API gives the same bits as the batched one
    one = reg.apply_transform(d_imgs[3], d_imgs[0], tfm, 0, sk.sitkNearestNeighbor)
    assert bool((one.tensor == outs[3].tensor).all())
    # oracle on the image and two of the masks (full size, seconds on the host cores)
    g = orc.geom_of(fixed)
    chain = [("dvf", dvf.array, g)]
    exp_img = orc.resample_scalar(moving.array, g, g, chain, 2, -1000.0)
    assert np.array_equal(engine.to_host(outs[0], pinned=False).array, exp_img)
    for k in (1, 20):
        exp = orc.resample_scalar(labels[k - 1].array, g, g, chain, 1, 0)
        assert np.array_equal(engine.to_host(outs[k], pinned=False).array, exp)
    # zero displacement -> identity for nearest-neighbour labels
    zero = sk.DisplacementFieldTransform(Image(np.zeros(SIZE[::-1] + (3,)), is_vector=True))
    ident = reg.apply_transform(d_imgs[5], d_imgs[0], zero, 0, sk.sitkNearestNeighbor)
    assert bool((ident.tensor == d_imgs[5].tensor).all())


def test_cfg2_whole_registration_matches_oracle(engine, pair):
    """BASELINE.json configs[1] end to end at full size: 3-level pyramid [4, 2, 1], 100 / 50 / 25 iterations -- pyramid Gaussians
    (sigma 4 / 2 / 1 mm), every Demons iteration with the reference's early stop, field re-gridding, composition and the recursive
    Gaussian between levels, the final warp -- against the oracle's restatement of deformable.py:31-306 (about a minute of host
    cores).  Asserts equal elapsed iterations per level, DVF within the north star's 1e-4 mm (in fact the same bits) and the
    registered image within 1e-5 relative."""
    from oracle import platipy_ref as ref

    fixed, moving = pair
    kw = dict(resolution_staging=[4, 2, 1], iteration_staging=[100, 50, 25])
    stats = []
    img_o, _, dvf_o = ref.fast_symmetric_forces_demons_registration(fixed, moving, level_stats=stats, **kw)
    img, tfm, dvf = reg.fast_symmetric_forces_demons_registration(engine.to_device(fixed), engine.to_device(moving), **kw)
    got_stats = reg.LAST_LEVEL_STATS[:]
    assert [s["elapsed_iterations"] for s in got_stats] == [s["elapsed_iterations"] for s in stats]
    for g, o in zip(got_stats, stats):
        assert abs(g["metric"] - o["metric"]) <= 1e-9 * abs(o["metric"])
    got = engine.to_host(dvf, pinned=False).array
    err = float(np.abs(got - dvf_o.array).max())
    print(f"cfg2 whole registration: elapsed {[s['elapsed_iterations'] for s in got_stats]}, max |DVF gpu - oracle| = {err:.3e} mm, "
          f"identical bits: {np.array_equal(got, dvf_o.array)}")
    assert err <= 1e-4
    gi = engine.to_host(img, pinned=False).array
    assert float(np.abs(gi - img_o.array).max()) <= 1e-5 * max(1.0, float(np.abs(img_o.array).max()))


def test_cfg4_size_atlas_pipeline_matches_oracle(engine):
    """BASELINE.json configs[3] at its size (256x256x160, spacing 1 x 1 x 1.5, 4 atlases x 5 structures, Demons [4, 2, 1] x [50, 50, 25],
    unweighted vote): Demons -> batched (bit-packed) label propagation -> vote counts -> finalisation -> process_probability_image on the GPU
    against the oracle pipeline run atlas by atlas on the host cores (about a minute).  The atlases are generated already aligned with
    the target (the linear step has its own functional tests; its optimiser has no oracle).  Fused probabilities and masks: same bits."""
    from oracle import platipy_ref as ref
    from platipy_b200 import multiatlas
    from platipy_b200.synth import synth_atlas_case

    size, sp, n_struct = (256, 256, 160), (1.0, 1.0, 1.5), 5
    target, _ = synth_atlas_case(size, n_struct, sp, seed=0, atlas_seed=None)
    names = [f"S{k}" for k in range(n_struct)]
    atlas_set = {}
    for a in range(4):
        ct, labs = synth_atlas_case(size, n_struct, sp, seed=0, atlas_seed=a, similarity=False)
        atlas_set[f"{a:03d}"] = dict({"CT Image": ct}, **dict(zip(names, labs)))
    dset = {"isotropic_resample": False, "resolution_staging": [4, 2, 1], "iteration_staging": [50, 50, 25], "smoothing_sigmas": [4, 2, 0],
            "ncores": 32, "default_value": -1000, "verbose": False}
    settings = {"deformable_registration_settings": dset,
                "label_fusion_settings": {"vote_type": "unweighted", "vote_params": None, "optimal_threshold": {}, "fusion": "vote"}}
    results, probs = multiatlas.run_segmentation(target, atlas_set, settings)
    kw = {k: v for k, v in dset.items() if k not in ("ncores", "verbose")}
    dirs = {}
    for a in sorted(atlas_set):
        _, tfm, _ = ref.fast_symmetric_forces_demons_registration(target, atlas_set[a]["CT Image"], **kw)
        d = {n: ref.apply_transform(atlas_set[a][n], target, tfm, 0, sk.sitkNearestNeighbor) for n in names}
        d["Weight Map"] = ref.compute_weight_map(target, target, "unweighted", None)
        dirs[a] = {"DIR": d}
    exp = ref.combine_labels(dirs, names)
    for n in names:
        assert np.array_equal(probs[n].array, exp[n].array), n
        assert np.array_equal(results[n].array, ref.process_probability_image(exp[n], 0.5).array), n
        assert results[n].array.sum() > 0
