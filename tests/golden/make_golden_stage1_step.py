"""Golden vector for ONE composed stage-1 training update: the body of Higashi.train_epoch for
`use_recon=True` (higashi/Higashi_wrapper.py:970-1010) on the UNMODIFIED reference model in eval mode (dropout off):

    loss_bce.backward(retain_graph)  -> main_norm  = ||grad(wstack[0].weight_list[0])||
    loss_mse.backward(retain_graph)  -> recon_norm = ||grad(wstack[0].weight_list[0])||
    ratio1 = max(beta * main_norm / recon_norm, 100 * median_total_sparsity_cell - 3)
    train_loss = alpha * loss_bce + ratio1 * loss_mse + contractive_loss_weight * contractive_loss ; backward ; Adam.step

Run by hand in the build container (needs /root/reference):  python tests/golden/make_golden_stage1_step.py
Writes tests/golden/stage1_step_d64.npz (the other fixtures are not touched)."""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
import make_golden as MG  # noqa: E402  (imports the reference through the shims)
import torch  # noqa: E402

RW = MG.RW


def main():
    d = 64
    ds = MG.tiny_dataset()
    rng = np.random.Generator(np.random.PCG64(2024))
    out = MG.dataset_arrays(ds)
    model, nei = MG.build_reference_model(ds, d, seed=8)
    model.eval()
    x = MG.positives(ds, 0, 18, rng)
    x[9:, 2] = np.minimum(x[9:, 2] + 1, ds.num_list[1])
    y = np.concatenate([np.ones((9, 1)), np.zeros((9, 1))]).astype(np.float32)
    w = (rng.random((18, 1)) + 0.5).astype(np.float32)
    ns = type("NS", (), {})()
    ns.higashi_model = model; ns.use_recon = True; ns.node_embedding_init = nei; ns.mode = "classification"
    ns.cell_ids = torch.arange(ds.n_cells).long()
    alpha, beta, median_sparsity, cw = 1.0, 1e-2, 0.02, 1e-3
    out.update({"before/" + k[3:]: v.copy() for k, v in MG.state_arrays(model).items()})
    # the optimiser of train_for_embeddings covers the model and the embedding (Higashi_wrapper.py:1340-1349)
    params = list(dict.fromkeys(list(model.parameters()) + list(nei.parameters())))
    params = [p for p in params if p.requires_grad]
    opt = torch.optim.Adam(params, lr=1e-3)
    w0 = nei.wstack[0].weight_list[0]
    pred, loss_bce, loss_mse = RW.Higashi.forward_batch_hyperedge(
        ns, torch.from_numpy(x), torch.from_numpy(w), np.zeros(len(x), dtype="int8"), [], torch.from_numpy(y), np.array([0]))
    opt.zero_grad(set_to_none=True)
    loss_bce.backward(retain_graph=True)
    main_norm = w0.grad.data.norm(2)
    opt.zero_grad(set_to_none=True)
    loss_mse.backward(retain_graph=True)
    recon_norm = w0.grad.data.norm(2)
    ratio = beta * main_norm / recon_norm
    ratio1 = max(ratio, 100 * median_sparsity - 3)
    contractive = 0.0
    for i in range(len(nei.wstack[0].weight_list)):
        contractive = contractive + torch.sum(nei.wstack[0].weight_list[i] ** 2)
        contractive = contractive + torch.sum(nei.wstack[0].reverse_weight_list[i] ** 2)
    train_loss = alpha * loss_bce + ratio1 * loss_mse + cw * contractive
    opt.zero_grad(set_to_none=True)
    train_loss.backward()
    opt.step()
    out["x"] = x.astype(np.int64); out["y"] = y; out["w"] = w
    out["cfg/alpha"] = np.array(alpha); out["cfg/beta"] = np.array(beta); out["cfg/median_sparsity"] = np.array(median_sparsity)
    out["cfg/contractive_weight"] = np.array(cw); out["cfg/lr"] = np.array(1e-3)
    out["loss_bce"] = np.array(float(loss_bce)); out["loss_mse"] = np.array(float(loss_mse))
    out["main_norm"] = np.array(float(main_norm)); out["recon_norm"] = np.array(float(recon_norm)); out["ratio1"] = np.array(float(ratio1))
    out.update({"after/" + k[3:]: v.copy() for k, v in MG.state_arrays(model).items()})
    np.savez_compressed(os.path.join(HERE, "stage1_step_d64.npz"), **out)
    moved = {k[6:]: float(np.abs(out[k] - out["before/" + k[6:]]).max()) for k in out if k.startswith("after/")}
    print("stage1_step_d64.npz: bce %.4f mse %.4f main_norm %.4g recon_norm %.4g ratio1 %.4g; %d tensors, %d moved"
          % (out["loss_bce"], out["loss_mse"], out["main_norm"], out["recon_norm"], out["ratio1"], len(moved), sum(v > 0 for v in moved.values())))


if __name__ == "__main__":
    main()
