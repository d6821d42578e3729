"""Bodies of GPU tests executed on the CPU against tests/fake_engine.py -- an Engine stand-in answered by the oracle -- so that the
HOST GLUE of the package (everything above the C ABI: run_segmentation, the fusion and label utilities, linear_registration's
plumbing, the generators, iterative atlas removal, the comparison metrics) is covered by the CPU suite.  It matters most for the
rows written after the round's last GPU minute (tests/test_gpu_zzz_session3.py), which would otherwise not have executed at all:
that comparison.py composes contours / distance maps / statistics the way the reference does (its golden numbers come out through
the product's own functions), that the patch-correlation vote, the correlation / Mattes metrics, the L-BFGS-B optimiser, the moments
initialiser, get_bone_mask and get_com pass the right things in the right order.  It says nothing about the kernels (tests/emu and
the GPU tests do) and the stand-in is reachable only through this file's monkeypatch."""
import importlib.util
import os

import pytest

from platipy_b200.engine import Engine
from tests.fake_engine import FakeEngine

HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.fixture()
def fake(emu, monkeypatch):
    eng = FakeEngine(emu)
    monkeypatch.setattr(Engine, "get", classmethod(lambda cls, device=None: eng))
    return eng


@pytest.fixture(scope="module")
def gpu_tests():
    spec = importlib.util.spec_from_file_location("gpu_session3_tests", os.path.join(HERE, "test_gpu_zzz_session3.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


@pytest.mark.parametrize("name", ["test_surface_dsc", "test_surface_metrics", "test_metrics_match_the_oracle_on_irregular_labels",
                                  "test_patch_correlation_weight_map", "test_linear_registration_correlation_metric", "test_get_bone_mask",
                                  "test_alignment_registration_with_moments_and_lbfgsb", "test_linear_registration_mattes_mutual_information",
                                  "test_get_com", "test_gpu_matches_rows3_golden"])
def test_gpu_test_body_on_the_fake_engine(fake, gpu_tests, name):
    getattr(gpu_tests, name)(fake)
    assert fake.calls, "the test body did not reach the engine"


@pytest.fixture(scope="module")
def generation_tests():
    spec = importlib.util.spec_from_file_location("gpu_generation_tests", os.path.join(HERE, "test_gpu_zz_generation.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


@pytest.mark.parametrize("name,args", [("test_distance_map_helpers_bit_exact", ()), ("test_field_generators_bit_exact", ()),
                                       ("test_radial_bend_bit_exact", (("z", "inf"),)), ("test_radial_bend_bit_exact", (("y", "post"),)),
                                       ("test_radial_bend_bit_exact", (("x", "right"),)), ("test_radial_bend_bit_exact", (False,)),
                                       ("test_generators_with_demons_in_between", ()), ("test_device_in_device_out_and_augmentation", ()),
                                       ("test_iterative_atlas_removal", ())])
def test_generation_and_iar_glue_on_the_fake_engine(fake, generation_tests, name, args):
    """The same for the generators and iterative atlas removal (these did pass on a B200; here they keep the host glue covered by
    the CPU suite)."""
    getattr(generation_tests, name)(fake, *args)
    assert fake.calls


def _load(name):
    spec = importlib.util.spec_from_file_location("gpu_" + name, os.path.join(HERE, name + ".py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


@pytest.mark.parametrize("module,name,args", [
    ("test_gpu_multiatlas", "test_run_segmentation_matches_oracle_pipeline", ("vote",)),
    ("test_gpu_multiatlas", "test_run_segmentation_matches_oracle_pipeline", ("staple",)),
    ("test_gpu_multiatlas", "test_run_segmentation_with_linear_prealignment", ()),
    ("test_gpu_label_utils", "test_label_to_roi_crop_and_paste", ()),
    ("test_gpu_label_utils", "test_correct_volume_overlap_bit_exact", ()),
    ("test_gpu_label_utils", "test_binary_closing_matches_scipy", ()),
    ("test_gpu_fusion", "test_weight_maps_match_oracle", ()),
    ("test_gpu_fusion", "test_block_weight_map_and_normalise_match_oracle", ()),
    ("test_gpu_fusion", "test_combine_labels_bit_exact", ()),
    ("test_gpu_fusion", "test_staple_matches_oracle", ()),
    ("test_gpu_morph", "test_process_probability_image_matches_oracle", ()),
    ("test_gpu_linear", "test_masks_and_argument_errors", ()),
    ("test_gpu_parity", "test_smooth_and_resample_parity_and_errors", ()),                                  # blur + resample as one engine call
    ("test_gpu_parity", "test_float32_resample_through_field_on_output_grid", ((1, 0, 0, 0, 1, 0, 0, 0, 1),)),  # apply_transform, one image per call
])
def test_atlas_pipeline_glue_on_the_fake_engine(fake, module, name, args):
    """run_segmentation (auto-crop, linear pre-alignment, label propagation, Demons, weight maps, vote / STAPLE exchange, paste back,
    post-processing) and the label utilities: GPU-verified flows whose host logic the CPU suite keeps covered this way."""
    getattr(_load(module), name)(fake, *args)
    assert fake.calls


def test_every_session3_gpu_test_is_covered(gpu_tests):
    names = sorted(n for n in dir(gpu_tests) if n.startswith("test_"))
    covered = test_gpu_test_body_on_the_fake_engine.pytestmark[0].args[1]
    assert names == sorted(covered)
