"""Training-side kernels: three... mean-head forward, loss and the full backward pass against the reference's autograd
(golden fixture grads_d64.npz: one eval-mode classification batch through the stage-2 model) and the stage-1 / stage-2
forward outputs of scorer_d{64,128}.npz."""
import numpy as np
import pytest
import torch

from tests._util import FixtureDataset, load

pytestmark = pytest.mark.gpu


def _train_engine(g, prefix, ds, cell_table=None):
    from higashi_b200.train_engine import TrainEngine
    sd = {k[len(prefix):]: v for k, v in g.items() if k.startswith(prefix)}
    return TrainEngine(sd, ds.num, ds.feats, ds.cell_feats, device=0, cell_table=cell_table), sd


@pytest.mark.parametrize("d", [64, 128])
def test_forward_stage2_on_hook(d):
    g = load("scorer_d%d.npz" % d)
    ds = FixtureDataset(g)
    eng, _ = _train_engine(g, "stage2/sd/", ds, cell_table=g["stage2/cell_table"])
    for chrom in (0, 1):
        tn = (g["stage2/c%d/indices" % chrom], g["stage2/c%d/values" % chrom], g["stage2/c%d/unique" % chrom])
        mean, _ = eng.forward(g["stage2/c%d/x" % chrom], chrom, tn, stage=2)
        np.testing.assert_allclose(mean, g["stage2/c%d/out" % chrom][:, 0], atol=1e-5, rtol=0)
    eng.close()


@pytest.mark.parametrize("d", [64, 128])
def test_forward_stage1(d):
    g = load("scorer_d%d.npz" % d)
    ds = FixtureDataset(g)
    eng, _ = _train_engine(g, "stage1/sd/", ds)
    for chrom in (0, 1):
        mean, _ = eng.forward(g["stage1/c%d/x" % chrom], chrom, None, stage=1)
        np.testing.assert_allclose(mean, g["stage1/c%d/out" % chrom][:, 0], atol=1e-5, rtol=0)
    eng.close()


def test_backward_matches_autograd():
    g = load("grads_d64.npz")
    ds = FixtureDataset(g)
    eng, sd = _train_engine(g, "sd/", ds)
    tn = (g["indices"], g["values"], g["unique"])
    mean, loss = eng.forward(g["x"], 0, tn, y=g["y"], w=g["w"], stage=2)
    np.testing.assert_allclose(mean, g["out"][:, 0], atol=1e-5, rtol=0)
    assert abs(loss - float(g["loss"])) < 1e-5
    eng.zero_grad()
    eng.backward()
    torch.cuda.synchronize()
    grads = {k: v.cpu().numpy() for k, v in eng.named_grads().items()}
    checked = 0
    for k in g:
        if not k.startswith("grad/"):
            continue
        name = k[len("grad/"):]
        ref = g[k].reshape(grads[name].shape)
        scale = max(1e-3, float(np.abs(ref).max()))
        err = float(np.abs(grads[name] - ref).max())
        assert err <= 2e-5 * max(1.0, scale / 1e-2) + 1e-6, (name, err, scale)
        checked += 1
    assert checked >= 33
    # parameters that receive no gradient in the reference stay at zero (other chromosome's projection, unused heads)
    for name in ("encode1.static_nn.wstack.2.weight_list.0", "encode1.static_nn.wstack.0.weight_list.0",
                 "pff_classifier_var.w_stack.0.weight", "extra_proba.w_stack.0.weight"):
        assert float(np.abs(grads[name]).max()) == 0.0, name
    eng.close()


def test_adam_step_on_flat_views_reduces_loss():
    """torch.optim.Adam on views of the flat parameter tensor: a few steps on one batch lower the loss."""
    g = load("grads_d64.npz")
    ds = FixtureDataset(g)
    eng, _ = _train_engine(g, "sd/", ds)
    tn = (g["indices"], g["values"], g["unique"])
    params = [torch.nn.Parameter(v) for v in eng.named_params().values()]
    for p_, gv in zip(params, eng.named_grads().values()):
        p_.grad = gv
    opt = torch.optim.Adam(params, lr=1e-3)
    losses = []
    for _ in range(6):
        _, loss = eng.forward(g["x"], 0, tn, y=g["y"], w=g["w"], stage=2)
        losses.append(loss)
        eng.zero_grad()
        eng.backward()
        opt.step()
    assert losses[-1] < losses[0] - 1e-3, losses
    eng.close()


def test_data_parallel_two_gpus():
    """NCCL path: scripts/dp_train_check.py under torchrun on 2 GPUs (skipped on a 1-GPU box; the exchange step itself is
    covered on CPU by tests/test_parallel_cpu.py::test_two_rank_gradient_exchange)."""
    import os
    import subprocess
    import sys
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr",
                          "127.0.0.1", "--master-port", "29577", os.path.join(root, "scripts", "dp_train_check.py")],
                         capture_output=True, text=True, timeout=600)
    assert out.returncode == 0 and "DP_TRAIN_OK world=2" in out.stdout, out.stdout[-2000:] + out.stderr[-2000:]


@pytest.mark.parametrize("mode", ["classification", "regression", "zinb", "rank"])
def test_loss_modes_forward_backward(mode):
    """forward_batch_hyperedge loss heads (Higashi_wrapper.py:877-912) + autograd gradients, all four loss modes."""
    from higashi_b200.train_engine import LOSS_MODES
    g = load("losses_d64.npz")
    ds = FixtureDataset(g)
    eng, _ = _train_engine(g, "s2/sd/", ds)
    eng.set_options(cell2weight=g["cell2weight"], rank_thres=1.0)
    tn = (g["s2/indices"], g["s2/values"], g["s2/unique"])
    x, y, w = g["s2/x"], g["s2/y"], g["s2/w"]
    mean, loss = eng.forward(x, 1, tn, y=y, w=w, stage=2, loss_mode=LOSS_MODES[mode])
    heads = g["s2/%s/heads" % mode]
    np.testing.assert_allclose(mean, heads[:, 0], atol=1e-5, rtol=0)
    if mode == "zinb":
        m, v, p, pred = eng.heads(len(x), zinb=True)
        np.testing.assert_allclose(v, heads[:, 1], atol=1e-5, rtol=0)
        np.testing.assert_allclose(p, heads[:, 2], atol=1e-5, rtol=0)
        np.testing.assert_allclose(pred, g["s2/zinb/pred"].reshape(-1), atol=1e-5, rtol=1e-5)
    assert abs(loss - float(g["s2/%s/loss" % mode])) < 2e-5 * max(1.0, abs(float(g["s2/%s/loss" % mode]))), (loss, float(g["s2/%s/loss" % mode]))
    eng.zero_grad()
    eng.backward()
    torch.cuda.synchronize()
    grads = {k: v.cpu().numpy() for k, v in eng.named_grads().items()}
    pre = "s2/%s/grad/" % mode
    n = 0
    for k in g:
        if not k.startswith(pre):
            continue
        name = k[len(pre):]
        if name not in grads:
            continue
        ref = g[k].reshape(grads[name].shape)
        scale = max(1e-3, float(np.abs(ref).max()))
        err = float(np.abs(grads[name] - ref).max())
        assert err <= 5e-5 * max(1.0, scale / 1e-2) + 2e-6, (mode, name, err, scale)
        n += 1
    assert n >= 33, n
    eng.close()


def test_stage1_cell_branch_and_reconstruction():
    """Stage 1 (train_for_embeddings): cell branch on-hook + auto-encoder reconstruction MSE over all cells; the two
    gradient sets of train_epoch's separate backward passes (Higashi_wrapper.py:970-983)."""
    from higashi_b200.train_engine import TrainEngine
    g = load("losses_d64.npz")
    ds = FixtureDataset(g)
    sd = {k[len("s1/sd/"):]: v for k, v in g.items() if k.startswith("s1/sd/")}
    eng = TrainEngine(sd, ds.num, ds.feats, ds.cell_feats, device=0, cell_on_hook=True)
    mean, loss = eng.forward(g["s1/x"], 0, None, y=g["s1/y"], w=g["s1/w"], stage=1)
    np.testing.assert_allclose(mean, g["s1/pred"].reshape(-1), atol=1e-5, rtol=0)
    assert abs(loss - float(g["s1/loss"])) < 1e-5
    eng.zero_grad(); eng.backward(); torch.cuda.synchronize()
    grads = {k: v.cpu().numpy().copy() for k, v in eng.named_grads().items()}
    n = 0
    for k in g:
        if k.startswith("s1/grad_main/") and k[len("s1/grad_main/"):] in grads:
            name = k[len("s1/grad_main/"):]
            ref = g[k].reshape(grads[name].shape)
            assert float(np.abs(grads[name] - ref).max()) <= 3e-5 * max(1.0, float(np.abs(ref).max()) / 1e-2) + 2e-6, name
            n += 1
    assert n >= 33 and "encode1.static_nn.wstack.0.weight_list.0" in grads
    assert float(np.abs(grads["encode1.static_nn.wstack.0.weight_list.0"]).max()) > 0
    # reconstruction loss: separate gradient set, applied with a scale
    rl, gn = eng.recon_loss()
    assert abs(rl - float(g["s1/mse"])) < 1e-4 * float(g["s1/mse"])
    ref_w = g["s1/grad_recon/encode1.static_nn.wstack.0.weight_list.0"]
    assert abs(gn - float(np.linalg.norm(ref_w))) < 1e-4 * float(np.linalg.norm(ref_w))
    eng.zero_grad(); eng.recon_apply(1.0); torch.cuda.synchronize()
    gr = {k: v.cpu().numpy() for k, v in eng.named_grads().items()}
    for short in ("weight_list", "reverse_weight_list", "bias_list", "recon_bias_list"):
        name = "encode1.static_nn.wstack.0.%s.0" % short
        ref = g["s1/grad_recon/" + name]
        assert float(np.abs(gr[name] - ref).max()) <= 2e-5 * max(1.0, float(np.abs(ref).max()) / 1e-2) + 1e-6, name
    eng.close()


def test_dropout_gradient_is_consistent():
    """Training-mode dropout: with a fixed seed the masks of forward and backward coincide, so the analytic gradient
    matches a central finite difference of the (dropped-out) loss along a random direction."""
    g = load("grads_d64.npz")
    ds = FixtureDataset(g)
    eng, _ = _train_engine(g, "sd/", ds)
    tn = (g["indices"], g["values"], g["unique"])
    eng.set_dropout(True, seed=1234)
    _, l_drop = eng.forward(g["x"], 0, tn, y=g["y"], w=g["w"], stage=2)
    eng.set_dropout(False)
    _, l_eval = eng.forward(g["x"], 0, tn, y=g["y"], w=g["w"], stage=2)
    assert abs(l_drop - l_eval) > 1e-4                      # dropout really changes the forward
    eng.set_dropout(True, seed=1234)
    eng.forward(g["x"], 0, tn, y=g["y"], w=g["w"], stage=2)
    eng.zero_grad(); eng.backward(); torch.cuda.synchronize()
    gen = torch.Generator(device="cuda").manual_seed(0)
    v = torch.randn(eng.total, device="cuda", generator=gen)
    v /= v.norm()
    analytic = float((eng.grads.double() * v.double()).sum())
    eps = 2e-2
    base = eng.params.clone()
    eng.params.copy_(base + eps * v); _, lp = eng.forward(g["x"], 0, tn, y=g["y"], w=g["w"], stage=2)
    eng.params.copy_(base - eps * v); _, lm = eng.forward(g["x"], 0, tn, y=g["y"], w=g["w"], stage=2)
    eng.params.copy_(base)
    numeric = (lp - lm) / (2 * eps)
    assert abs(numeric - analytic) < 0.05 * max(abs(analytic), 1e-3) + 2e-4, (numeric, analytic)
    eng.close()


def test_stage1_trainer_step_runs_and_learns():
    from higashi_b200 import parallel
    from higashi_b200.train_engine import TrainEngine
    g = load("losses_d64.npz")
    ds = FixtureDataset(g)
    sd = {k[len("s1/sd/"):]: v for k, v in g.items() if k.startswith("s1/sd/")}
    eng = TrainEngine(sd, ds.num, ds.feats, ds.cell_feats, device=0, cell_on_hook=True)
    tr = parallel.DataParallelTrainer(eng, lr=1e-3)
    losses, recon = [], []
    for _ in range(8):
        losses.append(tr.step(g["s1/x"], 0, None, g["s1/y"], g["s1/w"], stage=1, train_mode=False, beta=1e-2,
                              median_sparsity=0.02, contractive_weight=1e-3))
        recon.append(tr.recon_loss)
    assert losses[-1] < losses[0] and recon[-1] < recon[0], (losses, recon)
    eng.close()


def test_stage1_composed_update_matches_the_reference():
    """ONE composed stage-1 update (train_epoch with use_recon=True, Higashi_wrapper.py:970-1010: gradient norm
    of the main loss, gradient norm of the reconstruction loss, ratio1, contractive term, Adam step) against a golden minted from
    the UNMODIFIED reference in eval mode (tests/golden/make_golden_stage1_step.py): the losses, both norms, ratio1 and EVERY
    parameter after the step."""
    from higashi_b200 import parallel
    from higashi_b200.train_engine import TrainEngine
    g = load("stage1_step_d64.npz")
    ds = FixtureDataset(g)
    sd = {k[len("before/"):]: v for k, v in g.items() if k.startswith("before/")}
    eng = TrainEngine(sd, ds.num, ds.feats, ds.cell_feats, device=0, cell_on_hook=True)
    tr = parallel.DataParallelTrainer(eng, lr=float(g["cfg/lr"]))
    loss = tr.step(g["x"], 0, None, g["y"], g["w"], stage=1, train_mode=False, alpha=float(g["cfg/alpha"]), beta=float(g["cfg/beta"]),
                   median_sparsity=float(g["cfg/median_sparsity"]), contractive_weight=float(g["cfg/contractive_weight"]))
    assert abs(loss - float(g["loss_bce"])) < 2e-5 and abs(tr.recon_loss - float(g["loss_mse"])) < 2e-5, (loss, tr.recon_loss)
    torch.cuda.synchronize()
    got = {k: v.detach().cpu().numpy() for k, v in eng.named_params().items()}
    checked = moved = n_sat = n_moved = 0
    for k in g:
        if not k.startswith("after/"):
            continue
        name = k[len("after/"):]
        if name not in got:
            continue
        want, before = g[k], g["before/" + name]
        # Adam's first step moves an entry by lr * g / (|g| + 1e-8): a full step of lr wherever |g| >> 1e-8.  Those entries check the sign
        # and presence of every gradient and the composition (ratio1, contractive term): they must agree to 5e-5.  Entries whose
        # reference step is not saturated have a gradient of the order of 1e-8 (e.g. the key LayerNorm bias, which only acts through
        # the masked diagonal): there the step size is rounding noise of the gradient, so only its bound is checked.
        lr = float(g["cfg/lr"])
        gotv = got[name].reshape(want.shape)
        sat = np.abs(want - before) >= 0.95 * lr
        if sat.any():
            assert np.abs(gotv - want)[sat].max() <= 5e-5, (name, float(np.abs(gotv - want)[sat].max()))
        assert np.abs(gotv - before).max() <= 1.05 * lr, name
        if not (want != before).any():                                  # no gradient at all in the reference (grad = None): untouched here
            assert np.array_equal(gotv, before), name
        n_sat += int(sat.sum()); n_moved += int((want != before).sum())
        checked += 1
        moved += int(np.abs(want - before).max() > 0)
    assert n_sat >= 0.9 * n_moved, (n_sat, n_moved)
    assert checked >= 40 and moved >= 20, (checked, moved)
    eng.close()


def test_module_forward_is_differentiable_like_the_reference():
    """Hyper_SAGNN.forward of the mirror under torch autograd (the reference's train_epoch body: model(x, (chrom, to_neighs)),
    F.binary_cross_entropy_with_logits, loss.backward()): heads, loss and every parameter gradient against the reference's
    autograd on the golden batch (grads_d64.npz, eval mode)."""
    import torch.nn.functional as F
    from tests.test_host_mirror import build_mirror_model
    g = load("grads_d64.npz")
    ds = FixtureDataset(g)
    sd = {k[3:]: v for k, v in g.items() if k.startswith("sd/")}
    model = build_mirror_model(ds, 64, sd)
    model.eval()
    model.encode1.static_nn.off_hook([0])                      # stage >= 2: frozen cell table (Higashi_wrapper.py:1384)
    x = torch.from_numpy(np.asarray(g["x"]))
    tn = (g["indices"], g["values"], g["unique"])
    mean, var, proba = model(x, (np.zeros(len(x), dtype=np.int8), tn), chroms_in_batch=np.array([1]))
    assert mean.shape == (len(x), 1) and var.shape == mean.shape and proba.shape == mean.shape and mean.requires_grad
    heads = torch.cat([mean, var, proba], dim=1).detach().cpu().numpy()          # the reference returns the same three heads
    np.testing.assert_allclose(heads, g["out"], atol=1e-5, rtol=0)
    y = torch.from_numpy(np.asarray(g["y"], dtype=np.float32)).view(-1, 1).to(mean.device)
    w = torch.from_numpy(np.asarray(g["w"], dtype=np.float32)).view(-1, 1).to(mean.device)
    loss = F.binary_cross_entropy_with_logits(mean, y, weight=w)
    assert abs(float(loss.detach()) - float(g["loss"])) < 1e-5
    loss.backward()
    named = dict(model.named_parameters(remove_duplicate=False))
    checked = 0
    for k in g:
        if not k.startswith("grad/"):
            continue
        name = k[len("grad/"):]
        got = named[name].grad
        assert got is not None, name
        ref = g[k].reshape(tuple(got.shape))
        scale = max(1e-3, float(np.abs(ref).max()))
        err = float(np.abs(got.cpu().numpy() - ref).max())
        assert err <= 2e-5 * max(1.0, scale / 1e-2) + 1e-6, (name, err, scale)
        checked += 1
    assert checked >= 33
    assert named["encode1.static_nn.wstack.2.weight_list.0"].grad is None          # the other chromosome's branch: untouched
    assert named["pff_classifier_var.w_stack.0.weight"].grad is None               # heads that did not feed the loss
    # a second call after an optimizer step sees the updated parameters
    opt = torch.optim.SGD([p for p in model.parameters() if p.grad is not None], lr=1e-2)
    opt.step()
    mean2, _, _ = model(x, (np.zeros(len(x), dtype=np.int8), tn))
    loss2 = F.binary_cross_entropy_with_logits(mean2, y, weight=w)
    assert float(loss2.detach()) < float(loss.detach())
