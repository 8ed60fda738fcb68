"""B200: the kernels of the steps either side of the hot path (csrc/preproc.cu, called through the C-ABI via
clip_ebc_b200.{transforms,eval_utils,evaluate}) against the fixtures produced by the reference's own functions and
against the CPU oracle (oracle/eval_oracle.py) on the same seeded inputs.

Tolerances: these are fp32 kernels restating fp32 PyTorch arithmetic with a different summation order / FMA contraction,
so values agree to a few ulp -- 2e-5 absolute on O(1) pixel values, 1e-5 relative on sums. Integer results (output
sizes, tap ranges) are exact."""
import numpy as np
import pytest
import torch

from oracle import clip_ebc_oracle as O
from oracle import eval_oracle as E
from oracle.golden_cases import (CASES, RESIZE_DENSITY_CASES, SUB_X, SUB_Y, TRANSFORM_CASES, case_inputs, make_density,
                                 make_points, make_u8_image)

from . import parity

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("case", RESIZE_DENSITY_CASES, ids=[c["name"] for c in RESIZE_DENSITY_CASES])
def test_resize_density_map_matches_reference_fixture(case):
    from clip_ebc_b200 import resize_density_map

    gold = parity.load_golden(case["name"])
    x = make_density(case["shape"], case["seed"], case["zero_image"])
    y_cpu_in = resize_density_map(x, case["size"])            # CPU in -> CPU out (what the reference's callers hold)
    y_dev = resize_density_map(x.cuda(), case["size"])        # device in -> device out
    assert y_cpu_in.device.type == "cpu" and y_dev.is_cuda and torch.equal(y_dev.cpu(), y_cpu_in)
    assert tuple(y_cpu_in.shape) == gold["out"].shape
    scale = max(1.0, float(np.abs(gold["out"]).max()))
    assert np.abs(y_cpu_in.numpy() - gold["out"]).max() <= 1e-5 * scale
    if case["zero_image"] is not None:
        assert float(y_cpu_in.abs().max()) == 0.0             # 0/0 -> nan_to_num -> 0, exactly
    else:
        want = float(gold["out_sum"][0])
        assert abs(float(y_cpu_in.double().sum()) - want) <= 1e-5 * abs(want)


def test_resize_density_map_is_deterministic_at_full_size():
    """Full-size property (a 384 x 512 map of BASELINE configs[4] back to 3072 x 4096): the two-stage sums have a fixed
    order, so repeated calls are bit-identical, and out == bilinear * (sum(bilinear) / sum(x)) with the kernel's own sums."""
    from clip_ebc_b200 import ops

    x = make_density((384, 512), 5).cuda()
    a, sums = ops.resize_density_map(x, (3072, 4096), return_sums=True)
    b = ops.resize_density_map(x, (3072, 4096))
    assert torch.equal(a, b)
    ref = torch.nn.functional.interpolate(x[None, None], size=(3072, 4096), mode="bilinear")[0, 0]
    s_in, s_out = sums.cpu().tolist()
    assert abs(s_in - float(x.double().sum())) <= 1e-5 * s_in and abs(s_out - float(ref.double().sum())) <= 1e-5 * s_out
    assert ((a - ref * (s_out / s_in)).abs().max() <= 2e-5 * ref.abs().max() * (s_out / s_in)).item()


@pytest.mark.parametrize("case", TRANSFORM_CASES, ids=[c["name"] for c in TRANSFORM_CASES])
@pytest.mark.parametrize("src", ["uint8", "float"])
def test_transforms_match_reference_fixture(case, src):
    from clip_ebc_b200.transforms import Resize2Multiple, ZeroPad2Multiple, normalize, preprocess

    gold = parity.load_golden(case["name"])
    u8 = make_u8_image(case["shape"], case["seed"])
    h, w = case["shape"][1:]
    img = torch.from_numpy(u8).cuda()
    if src == "float":
        img = img.float() / 255.0
    for tag, t in (("resize", Resize2Multiple(case["window"], case["stride"])),
                   ("pad", ZeroPad2Multiple(case["window"], case["stride"]))):
        pts = make_points(case["n_points"], h, w, case["seed"] + 1000)
        out, lab = preprocess(img, t, label=pts)                      # fused: /255 -> transform -> Normalize
        assert tuple(out.shape) == tuple(gold[f"{tag}_shape"])
        assert np.abs(out[:, ::SUB_Y, ::SUB_X].cpu().numpy() - gold[f"{tag}_sub"]).max() <= 2e-5
        assert np.abs(out.double().sum(dim=(1, 2)).cpu().numpy() - gold[f"{tag}_chan_sum"]).max() <= 1e-5 * out[0].numel()
        assert np.array_equal(lab.numpy(), gold[f"{tag}_labels"])
        # the reference's own call sequence: transform(image, label) on the [0,1] float image, then Normalize
        if src == "float":
            o2, _ = t(img, make_points(case["n_points"], h, w, case["seed"] + 1000))
            o2 = normalize(o2)
            assert (o2 - out).abs().max().item() <= 2e-6
        # and the whole image against the CPU oracle
        ref = E.preprocess(u8, tag, case["window"], case["stride"])
        assert (out.cpu() - ref).abs().max().item() <= 2e-5


def test_transforms_identity_and_errors():
    from clip_ebc_b200.transforms import Resize2Multiple, ZeroPad2Multiple, preprocess

    img = torch.rand(3, 448, 672, device="cuda")
    lab = torch.zeros(0, 2)
    for t in (Resize2Multiple(224, 112), ZeroPad2Multiple(224, 112)):
        o, l = t(img, lab)
        assert o is img and l is lab                                   # already a multiple: returned untouched
    with pytest.raises(TypeError):
        preprocess(img, transforms=object())
    with pytest.raises(RuntimeError):
        preprocess(img.double())


def test_evaluate_loop_matches_oracle():
    """evaluate() over a tiny synthetic 'dataset' (three images of different sizes, preprocessed on the GPU with
    Resize2Multiple) against the CPU oracle of the whole chain: transforms -> sliding_window_predict -> count -> MAE/RMSE."""
    from clip_ebc_b200 import evaluate, get_model
    from clip_ebc_b200.eval_loop import predict_counts
    from clip_ebc_b200.transforms import Resize2Multiple, preprocess

    case = CASES[0]
    sd, tf, bins, anchors, reduction, _ = case_inputs(case)
    model = get_model("clip_vit_b_16", input_size=224, reduction=reduction, bins=bins, anchor_points=anchors,
                      prompt_type="word", num_vpt=32, vpt_drop=0.0, deep_vpt=True, text_features=tf)
    model.load_state_dict(sd, strict=True)
    model = model.to("cuda").eval()
    shapes, n_pts = [(3, 300, 350), (3, 230, 470), (3, 448, 448)], [5, 0, 11]
    t = Resize2Multiple(224, 224)
    loader, ref_images = [], []
    for i, (shape, n) in enumerate(zip(shapes, n_pts)):
        u8 = make_u8_image(shape, 200 + i)
        img = preprocess(torch.from_numpy(u8).cuda(), t)
        loader.append((img[None].cpu(), [make_points(n, shape[1], shape[2], 300 + i)], None))  # CPU batches like a DataLoader
        ref_images.append(E.preprocess(u8, "resize", 224, 224)[None])
    err = evaluate(model, loader, torch.device("cuda"), sliding_window=True, window_size=224, stride=224)
    ref_err, ref_counts = E.evaluate(lambda im: O.sliding_window_predict(im, sd, tf, anchors, reduction, 224, 224),
                                     ref_images, [[0] * n for n in n_pts])
    counts = predict_counts(model, [b[0] for b in loader], torch.device("cuda"), True, 224, 224)
    print(f"\n[evaluate] ours {err} counts {counts}\n           oracle {ref_err} counts {ref_counts}")
    for c, r in zip(counts, ref_counts):
        assert abs(c - r) <= parity.COUNT_REL * abs(r)
    assert abs(err["mae"] - ref_err["mae"]) <= parity.COUNT_REL * max(abs(r) for r in ref_counts)
    assert abs(err["rmse"] - ref_err["rmse"]) <= parity.COUNT_REL * max(abs(r) for r in ref_counts)


def test_cross_image_window_batching_is_bit_identical_to_per_image_calls():
    """clipebc_sliding_window_predict_batch: images of different sizes (shared-grid, off-grid and single-window ones) in
    one pass give exactly the maps and counts of separate sliding_window_predict calls, chunk boundaries inside images
    included (window_chunk=8 forces them)."""
    from clip_ebc_b200 import get_model, sliding_window_predict, sliding_window_predict_batch

    case = CASES[0]
    sd, tf, bins, anchors, reduction, _ = case_inputs(case)
    from oracle import weights

    for chunk in (0, 8):
        model = get_model("clip_vit_b_16", input_size=224, reduction=reduction, bins=bins, anchor_points=anchors,
                          prompt_type="word", num_vpt=32, vpt_drop=0.0, deep_vpt=True, text_features=tf, window_chunk=chunk)
        model.load_state_dict(sd, strict=True)
        model = model.to("cuda").eval()
        images = [weights.make_image((1, 3, h, w), seed=400 + i).cuda()
                  for i, (h, w) in enumerate([(448, 672), (300, 500), (224, 224), (448, 448), (230, 700)])]
        dens, cnt = sliding_window_predict_batch(model, images, 224, 112, return_device=True)
        assert len(dens) == len(images) and tuple(cnt.shape) == (len(images),)
        for i, im in enumerate(images):
            d1, c1 = sliding_window_predict(model, im, 224, 112, return_device=True, return_count=True)
            assert torch.equal(dens[i], d1), f"image {i} differs"
            assert torch.equal(cnt[i:i + 1], c1.reshape(1))
        dens_cpu, cnt_cpu = sliding_window_predict_batch(model, [im.cpu() for im in images], 224, 112)
        assert all(d.device.type == "cpu" for d in dens_cpu) and torch.equal(cnt_cpu, cnt.cpu())


def test_prefetched_host_images_give_the_same_counts_as_device_images():
    """predict_counts copies pinned host images on a side stream while the previous image is computed (`_prefetch_to_device`):
    the counts equal, bit for bit, those of the same images already resident on the device, in order, for images of
    different sizes (including one large enough to take several internal passes)."""
    from clip_ebc_b200 import get_model
    from clip_ebc_b200.eval_loop import predict_counts

    case = CASES[0]
    sd, tf, bins, anchors, reduction, _ = case_inputs(case)
    model = get_model("clip_vit_b_16", input_size=224, reduction=reduction, bins=bins, anchor_points=anchors,
                      prompt_type="word", num_vpt=32, vpt_drop=0.0, deep_vpt=True, text_features=tf)
    model.load_state_dict(sd, strict=True)
    model = model.to("cuda").eval()
    from oracle import weights

    sizes = [(448, 672), (1344, 1792), (224, 224), (672, 448), (1344, 1792), (448, 448)]
    host = [weights.make_image((1, 3, h, w), seed=700 + i).pin_memory() for i, (h, w) in enumerate(sizes)]
    dev = torch.device("cuda")
    on_device = predict_counts(model, [im.cuda() for im in host], dev, True, 224, 112)
    for _ in range(3):
        assert predict_counts(model, host, dev, True, 224, 112) == on_device
    assert predict_counts(model, (im[0] for im in host), dev, True, 224, 112) == on_device  # [3,H,W] items, a generator
