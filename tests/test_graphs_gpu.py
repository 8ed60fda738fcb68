"""CUDA-graph replay of model(x) (clip_ebc_b200/model.py: _forward_graphed) on a B200: bit-identical to the eager path,
launch accounting, cache invalidation."""
import pytest
import torch

from oracle import weights
from oracle.golden_cases import CASES, case_inputs

pytestmark = pytest.mark.gpu


def _model(case):
    from clip_ebc_b200 import get_model

    sd, tf, bins, anchors, reduction, _ = case_inputs(case)
    model = get_model("clip_vit_b_16", input_size=224, reduction=reduction, bins=bins, anchor_points=anchors,
                      prompt_type="word", num_vpt=case["num_vpt"], vpt_drop=0.0, deep_vpt=case["deep_vpt"], text_features=tf)
    model.load_state_dict(sd, strict=True)
    return model.to("cuda").eval()


def test_graph_replay_is_bit_identical_to_eager_and_counts_its_launches():
    from clip_ebc_b200 import _lib

    lib = _lib.load()
    case = [c for c in CASES if c["name"] == "c1_forward_r8_deep"][0]
    model = _model(case)
    xs = [weights.make_image((4, 3, 224, 224), seed=200 + i).cuda() for i in range(3)]
    model.use_cuda_graphs = False
    model(xs[0])                               # packs the weights (its conversion kernels are launches too)
    l0 = lib.clipebc_launch_count()
    eager = [model(x) for x in xs]
    per_call = (lib.clipebc_launch_count() - l0) // len(xs)
    assert per_call > 50
    model.use_cuda_graphs = True
    first = model(xs[0])                       # first sighting of the shape: eager
    assert not any("graph" in e for e in model._graph_cache.values())
    l1 = lib.clipebc_launch_count()
    second = model(xs[1])                      # capture + replay
    assert lib.clipebc_launch_count() - l1 == 2 * per_call   # the capture pass counts its launches, the replay reports them
    entry = next(iter(model._graph_cache.values()))
    assert entry["launches"] == per_call
    l2 = lib.clipebc_launch_count()
    third = model(xs[2])                       # replay only
    again = model(xs[0])
    assert lib.clipebc_launch_count() - l2 == 2 * per_call
    assert torch.equal(first, eager[0]) and torch.equal(second, eager[1]) and torch.equal(third, eager[2])
    assert torch.equal(again, eager[0])
    assert second.data_ptr() != third.data_ptr()  # results are copies, not views of the graph's static buffer
    # another batch size gets its own graph; train mode (logits, exp) too
    x8 = weights.make_image((8, 3, 224, 224), seed=210).cuda()
    a, b, c, d = model(x8), model(x8), model(x8), model(x8)  # the first call grows the workspaces (new library epoch)
    assert torch.equal(a, b) and torch.equal(b, c) and torch.equal(c, d)
    assert any("graph" in e and tuple(e["x"].shape) == (8, 3, 224, 224) for e in model._graph_cache.values())
    model.training = True
    outs = [model(xs[0]) for _ in range(3)]
    model.training = False
    assert all(torch.equal(o[0], outs[0][0]) and torch.equal(o[1], outs[0][1]) for o in outs)
    assert torch.equal(outs[0][1], eager[0])


def test_graph_cache_follows_the_weights():
    """New weights re-pack the model and drop the captured graphs."""
    case = [c for c in CASES if c["name"] == "c1_forward_r8_deep"][0]
    model = _model(case)
    x = weights.make_image((2, 3, 224, 224), seed=220).cuda()
    base = [model(x) for _ in range(3)][-1]
    sd = weights.make_state_dict(21, num_vpt=32, deep_vpt=True, variant="stress")
    model.load_state_dict(sd, strict=True)
    outs = [model(x) for _ in range(3)]
    model.use_cuda_graphs = False
    assert torch.equal(outs[-1], model(x)) and not torch.equal(outs[-1], base)


def test_capture_does_not_disturb_other_threads():
    """The capture uses capture_error_mode="thread_local": CUDA calls of OTHER threads that are illegal during a "global"
    capture (allocations, event queries -- a DataLoader's pin-memory thread does both) keep working while this thread
    captures, and the capture itself succeeds."""
    import threading

    case = [c for c in CASES if c["name"] == "c1_forward_r8_deep"][0]
    model = _model(case)
    x = weights.make_image((2, 3, 224, 224), seed=250).cuda()
    ref = model(x)                      # first sighting of the shape: eager
    stop, errors, n_ok = threading.Event(), [], [0]

    def noisy():
        try:
            while not stop.is_set():
                t = torch.empty(1 << 20, device="cuda")      # cudaMalloc through the caching allocator on fresh sizes
                ev = torch.cuda.Event()
                ev.record()
                ev.query()
                h = torch.empty(1 << 16).pin_memory()        # cudaHostAlloc
                del t, h
                torch.cuda.empty_cache()
                n_ok[0] += 1
        except Exception as e:  # pragma: no cover
            errors.append(e)

    th = threading.Thread(target=noisy, daemon=True)
    th.start()
    try:
        outs = [model(x) for _ in range(3)]                  # capture + replays while the other thread hammers the runtime
    finally:
        stop.set()
        th.join(timeout=30)
    assert not errors, errors
    assert n_ok[0] > 0
    assert any("graph" in e for e in model._graph_cache.values()), "the capture was abandoned"
    assert all(torch.equal(o, ref) for o in outs)


def test_graphs_do_not_outlive_a_workspace_reallocation():
    """A larger batch grows the library's workspaces (free + malloc): graphs captured before hold the old addresses and
    must not be replayed -- the cache key carries the library's epoch, which every (re)allocation bumps."""
    case = [c for c in CASES if c["name"] == "c1_forward_r8_deep"][0]
    model = _model(case)
    x_small = weights.make_image((2, 3, 224, 224), seed=240).cuda()
    small = [model(x_small) for _ in range(3)][-1]            # eager, capture, replay
    assert any("graph" in e for e in model._graph_cache.values())
    x_big = weights.make_image((40, 3, 224, 224), seed=241).cuda()
    big = [model(x_big) for _ in range(4)][-1]                # grows every workspace, then captures its own graph
    again = [model(x_small) for _ in range(3)]
    model.use_cuda_graphs = False
    assert torch.equal(model(x_small), small) and torch.equal(model(x_big), big)
    assert all(torch.equal(a, small) for a in again)


def test_graph_path_steps_aside_for_profiling_and_outer_captures():
    from clip_ebc_b200 import _lib

    lib = _lib.load()
    case = [c for c in CASES if c["name"] == "c1_forward_r8_deep"][0]
    model = _model(case)
    x = weights.make_image((2, 3, 224, 224), seed=230).cuda()
    ref = [model(x) for _ in range(3)][-1]
    lib.clipebc_profile_enable(1)
    try:
        assert not model._graphs_usable()
        assert torch.equal(model(x), ref)
    finally:
        lib.clipebc_profile_enable(0)
    # a caller's own capture sees the library's launches directly (no nested replay)
    g = torch.cuda.CUDAGraph()
    sx = x.clone()
    with torch.cuda.graph(g):
        out = model(sx)
    g.replay()
    torch.cuda.synchronize()
    assert torch.equal(out, ref)
