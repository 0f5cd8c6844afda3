"""Multi-PatchGAN Gram head on a B200: the C-ABI kernels against the fp64 oracle, the reference's own
golden outputs, and the drop-in modules against the fp32 torch port on the same GPU and weights."""
import os

import numpy as np
import pytest
import torch

from conftest import ROOT
from oracle import patchgan_fp64 as O
from oracle.torch_port import patchgan_forward, patchgan_multiscale_forward

pytestmark = pytest.mark.gpu
GOLDEN = os.path.join(ROOT, "tests", "golden", "patchgan_head.npz")


def rel(a, b):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    return float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-30))


def npf(t):
    return t.detach().float().cpu().numpy()


def head_modules(params, e, nc, d):
    a1 = torch.nn.MultiheadAttention(e, 8)
    a2 = torch.nn.MultiheadAttention(e, 8)
    cl = torch.nn.Linear(e, nc)
    fp = torch.nn.Linear(d * d, e)
    sd = {k: torch.from_numpy(np.asarray(v)) for k, v in params.items()}
    for name, mod in (("attention_per_layer", a1), ("attention_per_patch", a2), ("classifier", cl), ("feature_projection", fp)):
        mod.load_state_dict({k[len(name) + 1:]: v for k, v in sd.items() if k.startswith(name + ".")})
    return a1.cuda(), a2.cuda(), cl.cuda(), fp.cuda()


def run_head(maps, a1, a2, cl, fp, ln_input=True):
    from heuristique_style_transfer_code_b200 import ops
    gram, norms = ops.patch_gram(maps, ln_input=ln_input)
    L, b, dd = gram.shape
    feat = ops.gemm_f32(gram.view(L * b, dd), fp.weight.detach().t(), fp.bias.detach()).view(L, b, -1)
    emb, out = ops.patch_attention(feat, a1, a2, cl)
    return gram, norms, feat, emb, out


def test_head_reproduces_the_reference_golden_outputs():
    """The committed vectors are outputs of the unmodified reference class: x_proj maps in, gram_norms / embeddings /
    output out. fp32 kernels against fp32 reference: 1e-5 normwise."""
    z = np.load(GOLDEN)
    n = sum(1 for k in z.files if k.startswith("x_proj_"))
    maps = [torch.from_numpy(z[f"x_proj_{i}"]).cuda() for i in range(n)]
    params = {k: z["param/" + k] for k in O.HEAD_KEYS}
    a1, a2, cl, fp = head_modules(params, 16, 5, 16)
    gram, norms, feat, emb, out = run_head(maps, a1, a2, cl, fp)
    assert rel(npf(norms), z["gram_norms"]) < 1e-5
    assert rel(npf(emb), z["embeddings"]) < 1e-5
    assert rel(npf(out), z["output"]) < 1e-5
    assert (npf(out).argmax(1) == z["output"].argmax(1)).all()


@pytest.mark.parametrize("d,e,shapes,b", [
    (64, 64, [(112, 112), (56, 56), (28, 28), (14, 14), (13, 13), (12, 12)], 3),     # patch 70 at 224x224
    (64, 64, [(112, 112), (56, 56), (55, 55), (54, 54)], 2),                         # patch 10 at 224x224
    (32, 32, [(24, 17), (5, 9), (3, 3), (2, 1), (1, 1)], 4),                          # tiny / ragged maps
    (128, 128, [(20, 20), (7, 7)], 2),                                               # the largest D and ndf accepted
    (24, 40, [(9, 31)] * 8, 2)])                                                     # eight layers, D not a multiple of 32
@pytest.mark.parametrize("ln_input", [True, False])
def test_kernels_against_the_fp64_oracle(d, e, shapes, b, ln_input):
    torch.manual_seed(d + len(shapes))
    maps = [torch.randn(b, d, h, w, device="cuda") * (1.0 + i) + 0.3 * i for i, (h, w) in enumerate(shapes)]
    a1 = torch.nn.MultiheadAttention(e, 8).cuda()
    a2 = torch.nn.MultiheadAttention(e, 8).cuda()
    cl = torch.nn.Linear(e, 7).cuda()
    fp = torch.nn.Linear(d * d, e).cuda()
    with torch.no_grad():
        for a in (a1, a2):
            a.in_proj_bias.normal_(0, 0.2)
            a.out_proj.bias.normal_(0, 0.2)
    params = {}
    for name, mod in (("attention_per_layer", a1), ("attention_per_patch", a2), ("classifier", cl), ("feature_projection", fp)):
        params.update({f"{name}.{k}": npf(v) for k, v in mod.state_dict().items()})
    want = O.patch_head([npf(m) for m in maps], params, heads=8, ln_input=ln_input)
    gram, norms, feat, emb, out = run_head(maps, a1, a2, cl, fp, ln_input)
    assert rel(npf(gram), want["grams"].reshape(len(shapes), b, -1)) < 2e-6
    assert rel(npf(norms), want["gram_norms"]) < 2e-6
    assert rel(npf(feat), want["projected"]) < 5e-5            # split-bf16 tensor-core GEMM (~1e-5)
    assert rel(npf(emb), want["embeddings"]) < 5e-5
    assert rel(npf(out), want["output"]) < 5e-5


def test_strided_and_channels_last_maps_give_the_same_gram():
    from heuristique_style_transfer_code_b200 import ops
    torch.manual_seed(5)
    x = torch.randn(3, 64, 14, 14, device="cuda")
    g0, n0 = ops.patch_gram([x])
    g1, n1 = ops.patch_gram([x.contiguous(memory_format=torch.channels_last)])
    big = torch.randn(3, 64, 16, 20, device="cuda")
    big[:, :, 1:15, 3:17] = x
    g2, n2 = ops.patch_gram([big[:, :, 1:15, 3:17]])
    for g, n in ((g1, n1), (g2, n2)):
        assert rel(npf(g), npf(g0)) < 1e-6 and rel(npf(n), npf(n0)) < 1e-6


def test_unsupported_shapes_are_refused():
    from heuristique_style_transfer_code_b200 import ops
    from heuristique_style_transfer_code_b200._lib import GramHeadError
    with pytest.raises(GramHeadError):
        ops.patch_gram([torch.randn(1, 160, 4, 4, device="cuda")])            # D > 128
    with pytest.raises(GramHeadError):
        ops.patch_gram([torch.randn(1, 8, 4, 4, device="cuda")] * 9)          # more than 8 layers
    with pytest.raises(GramHeadError):
        ops.patch_gram([torch.randn(1, 8, 4, 4)])                             # CPU tensor: no CPU path


@pytest.mark.parametrize("norm,patch", [("batch", 70), ("instance", 10), ("batch", 150)])
def test_discriminator_matches_the_torch_port(norm, patch):
    from heuristique_style_transfer_code_b200.patchgan import VariablePatchesNLayerDiscriminator_test
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.manual_seed(patch)
    m = VariablePatchesNLayerDiscriminator_test(ndf=64, norm=norm, patch_size=patch, num_classes=4, gram_matrix_dim=64).cuda().eval()
    x = torch.randn(4, 3, 224, 224, device="cuda")
    with torch.no_grad():
        emb, out = m(x)
        pe, po, pn = patchgan_forward(m, x)
    assert emb.shape == (4, 64) and out.shape == (4, 4)
    assert rel(npf(emb), npf(pe)) < 1e-4 and rel(npf(out), npf(po)) < 1e-4
    assert rel(npf(torch.stack(m.get_gram_norms())), npf(torch.stack(pn))) < 1e-5
    assert len(m.get_gram_norms()) == len(m.projection_layers) and m.get_gram_norms()[0].shape == (4,)
    assert (out.argmax(1) == po.argmax(1)).all()


def test_multiscale_matches_the_torch_port_and_refuses_grad_mode():
    from heuristique_style_transfer_code_b200.patchgan import MultiScaleDiscriminator_test
    from heuristique_style_transfer_code_b200._lib import GramHeadError
    torch.manual_seed(2)
    m = MultiScaleDiscriminator_test(ndf=64, norm='batch', num_classes=4, gram_matrix_dim=64).cuda().eval()
    x = torch.randn(2, 3, 224, 224, device="cuda")
    with torch.no_grad():
        emb, out = m(x)
        pe, po = patchgan_multiscale_forward(m, x)
    assert rel(npf(emb), npf(pe)) < 1e-4 and rel(npf(out), npf(po)) < 1e-4
    assert len(m.get_gram_norms()) == sum(len(d.projection_layers) for d in m.scale_discriminators.values())
    # A forward that needs autograd history (upstream: style_transfer_patches, functions_Multi_PatchGAN.py:272-287, which
    # backpropagates through the head to a noise image) is handed to the reference's OWN forward on this module's
    # parameters; without a copy of the reference it is refused instead of returning tensors without history.
    from heuristique_style_transfer_code_b200 import _reference
    if _reference.reference_root(os.path.join("Models", "Models_Multi_PatchGAN.py")) is not None:
        xg = x.clone().requires_grad_(True)
        e2, o2 = m(xg)
        assert e2.requires_grad and o2.requires_grad
        assert rel(npf(e2), npf(emb)) < 1e-4 and rel(npf(o2), npf(out)) < 1e-4
        e2.square().sum().backward()
        assert xg.grad is not None and bool(torch.isfinite(xg.grad).all()) and float(xg.grad.abs().sum()) > 0
    saved_root, saved_loaded = _reference.reference_root, dict(_reference._LOADED)
    try:
        _reference.reference_root = lambda rel_path: None
        _reference._LOADED.clear()
        with pytest.raises(GramHeadError, match="inference-only"):
            m(x)
    finally:
        _reference.reference_root = saved_root
        _reference._LOADED.update(saved_loaded)


def test_nan_inputs_take_the_reference_replacement_path(capsys):
    """A NaN pixel: the reference prints and zeroes NaNs layer by layer (:186-196); the drop-in detects the poisoned
    result, reruns with those checks and must land on the same numbers and messages as the port."""
    from heuristique_style_transfer_code_b200.patchgan import VariablePatchesNLayerDiscriminator_test
    torch.manual_seed(9)
    m = VariablePatchesNLayerDiscriminator_test(ndf=32, norm='batch', patch_size=30, num_classes=3, gram_matrix_dim=32).cuda().eval()
    x = torch.randn(2, 3, 64, 64, device="cuda")
    x[1, 0, 5, 7] = float("nan")
    with torch.no_grad():
        emb, out = m(x)
        ours_msgs = capsys.readouterr().out
        pe, po, _ = patchgan_forward(m, x)
        port_msgs = capsys.readouterr().out
    assert "NaN detected after layer 0" in ours_msgs and ours_msgs == port_msgs
    assert torch.isfinite(emb).all() and torch.isfinite(out).all()
    assert rel(npf(emb), npf(pe)) < 1e-4 and rel(npf(out), npf(po)) < 1e-4
