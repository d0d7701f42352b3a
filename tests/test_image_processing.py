"""ViLT image pre-processing: the oracle is pinned against Pillow itself (CPU); the CUDA kernels are
compared bit for bit with the oracle (GPU)."""
import numpy as np
import pytest
import torch

from oracle import image_oracle as IO

SHAPES = [(480, 640), (333, 500), (1000, 700), (224, 224), (97, 131), (384, 384), (2000, 1500), (50, 400), (600, 35), (31, 33)]


def _images(seed=0, shapes=SHAPES):
    rng = np.random.default_rng(seed)
    out = []
    for n, (h, w) in enumerate(shapes):
        if n % 3 == 0:  # smooth content as well as noise
            yy, xx = np.mgrid[0:h, 0:w]
            im = np.stack([(yy * 255 / max(h - 1, 1)), (xx * 255 / max(w - 1, 1)), ((yy + xx) % 256)], -1).astype(np.uint8)
        else:
            im = rng.integers(0, 256, (h, w, 3), dtype=np.uint8)
        out.append(im)
    return out


def test_output_size_rule_matches_hf_formula():
    from vault_b200.image_processing import resize_output_size

    assert IO.output_size(480, 640) == (384, 512) and IO.output_size(224, 224) == (384, 384)
    assert IO.output_size(50, 400) == (64, 608) and IO.output_size(400, 50) == (608, 64)  # int(1333/800*384) = 639 -> floor32 = 608
    rng = np.random.default_rng(1)
    for _ in range(500):
        h, w = int(rng.integers(20, 3000)), int(rng.integers(20, 3000))
        assert resize_output_size(h, w) == IO.output_size(h, w)
        ho, wo = IO.output_size(h, w)
        assert ho % 32 == 0 and wo % 32 == 0 and max(ho, wo) <= 608


def test_numpy_restatement_of_pillow_resample_is_bit_exact():
    from PIL import Image

    for im in _images():
        ho, wo = IO.output_size(*im.shape[:2])
        if min(ho, wo) == 0:
            continue
        want = np.array(Image.fromarray(im).resize((wo, ho), resample=Image.BICUBIC, reducing_gap=None))
        assert np.array_equal(IO.pillow_resample_u8(im, (ho, wo)), want), im.shape


def test_host_tap_tables_match_the_oracle_and_plan_is_consistent():
    from vault_b200.image_processing import ViltImageProcessorB200, pillow_bicubic_taps

    for i, o in ((640, 512), (333, 384), (97, 384), (2000, 512), (384, 384)):
        bounds, coefs, ksize = pillow_bicubic_taps(i, o)
        ref = IO._coeffs(i, o)
        assert all(bounds[x, 0] == ref[x][0] and bounds[x, 1] == len(ref[x][1]) and np.array_equal(coefs[x, :len(ref[x][1])], ref[x][1]) for x in range(o))
        assert int(coefs.sum(1).min()) >= (1 << 22) - ksize and int(coefs.sum(1).max()) <= (1 << 22) + ksize  # taps sum to one
    p = ViltImageProcessorB200(device="cpu").plan([(480, 640), (333, 500), (480, 640)])
    d = p["descs"]
    assert (p["Hmax"], p["Wmax"]) == (384, 576) and d["coef_h"][0] == d["coef_h"][2] and d["src_off"][1] == 480 * 640 * 3
    assert d.dtype.itemsize == 56 and p["src_bytes"] == (480 * 640 * 2 + 333 * 500) * 3
    with pytest.raises(RuntimeError, match="CUDA only"):
        ViltImageProcessorB200(device="cpu")([np.zeros((32, 32, 3), np.uint8)])


@pytest.mark.gpu
def test_gpu_preprocessing_is_bit_identical_to_the_reference_pipeline():
    from vault_b200.image_processing import ViltImageProcessorB200

    shapes = [s for s in SHAPES if min(IO.output_size(*s)) > 0]
    ims = _images(3, shapes)
    pv, pm, _ = IO.reference_preprocess(ims)
    out = ViltImageProcessorB200(device="cuda:0")(ims)
    assert out["pixel_values"].dtype == torch.float32 and out["pixel_mask"].dtype == torch.int64
    assert np.array_equal(out["pixel_mask"].cpu().numpy(), pm)
    assert np.array_equal(out["pixel_values"].cpu().numpy(), pv)  # bit-exact: integer resize + table lookup
    # a uniform batch (the TWITTER shape) and PIL / torch inputs
    from PIL import Image
    same = _images(4, [(300, 300)] * 5)
    pv2, pm2, _ = IO.reference_preprocess(same)
    out2 = ViltImageProcessorB200(device="cuda:0")([Image.fromarray(same[0]), torch.from_numpy(same[1])] + same[2:])
    assert np.array_equal(out2["pixel_values"].cpu().numpy(), pv2) and tuple(out2["pixel_values"].shape) == (5, 3, 384, 384)
