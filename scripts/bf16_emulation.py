"""Design study (CPU, numpy): error of the bf16 tensor-core pipeline of K2 against the fp32 hoisted math.
Rounding points emulated: Q/K/V token tables, ctx, LN(y), h, u operands and the GEMM weights in bf16;
accumulation, LayerNorm statistics, softmax, residual, square and heads in fp32.
Usage: python scripts/bf16_emulation.py"""
import sys, os
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from higashi_b200 import synthetic as S
from oracle import prep as OP, scorer as OS, impute as OI


def bf(x):
    return torch.as_tensor(x, dtype=torch.float32).to(torch.bfloat16).to(torch.float32).numpy()


def ln_stats(x):
    mu = x.mean(-1, keepdims=True)
    var = ((x - mu) ** 2).mean(-1, keepdims=True)
    return (x - mu) / np.sqrt(var + 1e-5)


def swish(x):
    return x / (1 + np.exp(-x))


def hoisted_scores(sd, tables, ds, cell, nbr, nbr_w, pairs_by_chrom, opt):
    """opt: dict of booleans selecting bf16 rounding points."""
    g = lambda k: np.asarray(sd[k], dtype=np.float32)
    r = (lambda x: bf(x)) if True else None
    A = "encode1.mul_head_attn."
    d = tables[0].shape[1]
    Wq, Wk, Wv, fc1, fc2 = g(A + "w_qs.weight"), g(A + "w_ks.weight"), g(A + "w_vs.weight"), g(A + "fc1.w_stack.0.weight"), g(A + "fc2.w_stack.0.weight")
    lnp = lambda n: (g(n + ".weight"), g(n + ".bias"))
    g1, b1 = lnp(A + "layer_norm1"); g2, b2 = lnp(A + "layer_norm2"); g3, b3 = lnp(A + "layer_norm3")
    gN, bN = lnp("encode1.pff_n1.layer_norm"); gL1, bL1 = lnp("layer_norm1"); gL2, bL2 = lnp("layer_norm2")
    W0, c0 = g("encode1.pff_n1.w_stack.0.weight")[:, :, 0], g("encode1.pff_n1.w_stack.0.bias")
    W1, c1 = g("encode1.pff_n1.w_stack.1.weight")[:, :, 0], g("encode1.pff_n1.w_stack.1.bias")
    C0, cb0 = g("pff_classifier.w_stack.0.weight")[:, :, 0], g("pff_classifier.w_stack.0.bias")
    C1, cb1 = g("pff_classifier.w_stack.1.weight")[:, :, 0], g("pff_classifier.w_stack.1.bias")
    G, gb = g("encode1.dynamic_nn.nn.weight"), g("encode1.dynamic_nn.nn.bias")
    rw = bf if opt["w"] else (lambda x: x)
    ra = bf if opt["act"] else (lambda x: x)
    rqk = bf if opt["qk"] else (lambda x: x)
    rv = bf if opt["v"] else (lambda x: x)
    out = []
    Ec = tables[0][cell - 1].numpy()
    Qc = (ln_stats(Ec[None]) * g1 + b1) @ Wq.T; Kc = (ln_stats(Ec[None]) * g2 + b2) @ Wk.T
    Vc = (ln_stats(Ec[None]) * g3 + b3) @ Wv.T
    Sc = ln_stats(Vc @ fc2.T) * gL2 + bL2
    W0f = W0 * gN[None, :]; c0f = c0 + W0 @ bN
    for j, pairs in pairs_by_chrom.items():
        E = tables[j + 1].numpy()
        if nbr is None:
            adj = OP.prep_one(lambda c: ds.csr.matrix(j, c - 1), 0, [cell], None)
        else:
            adj = OP.prep_one(lambda c: ds.csr.matrix(j, c - 1), 0, list(nbr[cell - 1]), list(nbr_w[cell - 1]))
        neigh = adj @ E
        dyn = swish(np.concatenate([neigh, E], -1) @ G.T + gb)
        Q = (ln_stats(dyn) * g1 + b1) @ Wq.T; K = (ln_stats(dyn) * g2 + b2) @ Wk.T
        V = (ln_stats(E) * g3 + b3) @ Wv.T
        Sb = ln_stats(V @ fc2.T) * gL2 + bL2
        qcK = (Qc.reshape(1, 8, 16) * K.reshape(-1, 8, 16)).sum(-1) * 0.25
        kcQ = (Q.reshape(-1, 8, 16) * Kc.reshape(1, 8, 16)).sum(-1) * 0.25
        Qr, Kr, Vr, Vcr = rqk(Q), rqk(K), rv(V), rv(Vc)
        bi, bj = pairs[:, 0], pairs[:, 1]
        s12 = (Qr[bi].reshape(-1, 8, 16) * Kr[bj].reshape(-1, 8, 16)).sum(-1) * 0.25
        s21 = (Qr[bj].reshape(-1, 8, 16) * Kr[bi].reshape(-1, 8, 16)).sum(-1) * 0.25

        def row(la, lb):
            mx = np.maximum(0, np.maximum(la, lb))
            e0, ea, eb = np.exp(-mx), np.exp(la - mx), np.exp(lb - mx)
            Z = e0 + ea + eb
            pa, pb = ea / Z, eb / Z
            den = pa + pb + 1e-13
            return pa / den, pb / den
        w01, w02 = row(qcK[bi], qcK[bj]); w10, w12 = row(kcQ[bi], s12); w20, w21 = row(kcQ[bj], s21)
        rep = lambda w: np.repeat(w, 16, axis=-1)
        ctx = [rep(w01) * Vr[bi] + rep(w02) * Vr[bj], rep(w10) * Vcr + rep(w12) * Vr[bj], rep(w20) * Vcr + rep(w21) * Vr[bi]]
        Ss = [np.broadcast_to(Sc, (len(bi), d)), Sb[bi], Sb[bj]]
        o = 0
        for t in range(3):
            y = ra(ctx[t]) @ rw(fc1).T
            x = ra(ln_stats(y)) @ rw(W0f).T + c0f
            h = swish(x) if not opt["tanh"] else x * (0.5 + 0.5 * np.tanh(0.5 * x) * (1 + opt["tanh"] * np.random.default_rng(0).uniform(-1, 1, x.shape)))
            z = y + ra(h) @ rw(W1).T + c1
            u = (ln_stats(z) * gL1 + bL1 - Ss[t]) ** 2
            c = swish(ra(u) @ rw(C0).T + cb0)
            o = o + (c @ C1.T + cb1)
        out.append((o / 3.0).reshape(-1))
    return np.concatenate(out)


def main():
    for d in (64, 128):
        ds = S.make_dataset("C1", n_cells=24, bins=[250, 60], seed=3)
        sd, feats = S.random_state_dict(ds, d, seed=7 + d)
        tsd = {k: torch.from_numpy(np.asarray(v)) for k, v in sd.items()}
        with torch.no_grad():
            tables = OS.embed_tables(tsd, feats)
        rng = np.random.Generator(np.random.PCG64(9))
        nbr, nbr_w = OP.cell_neighbors(rng.standard_normal((ds.n_cells, 16)).astype(np.float32), 5)
        nbr_w = nbr_w.astype(np.float32)
        chroms = [0, 1]
        big, bc, slices, coords = OI.enumerate_pairs(ds.num_list, chroms, 1, int(1e5))
        cell = 5
        ref = OI.impute_cells(tsd, tables, ds.cell_feats, ds.num_list, lambda c, cl: ds.csr.matrix(c, cl - 1), [cell - 1], chroms, big, 0, nbr, nbr_w, "sigmoid")[0]
        # pair bias
        x = big.copy(); x[:, 0] = cell
        with torch.no_grad():
            _, dp2, _ = OS.distance_side_channel(torch.from_numpy(x), tsd, ds.cell_feats, True)
        dp2 = dp2.numpy().reshape(-1)
        sig = lambda v: 1 / (1 + np.exp(-v))
        for name, opt in [("fp32 hoisted", dict(w=0, act=0, qk=0, v=0, tanh=0)),
                          ("bf16 GEMM operands only", dict(w=1, act=1, qk=0, v=0, tanh=0)),
                          ("+ V table bf16", dict(w=1, act=1, qk=0, v=1, tanh=0)),
                          ("+ Q,K tables bf16", dict(w=1, act=1, qk=1, v=1, tanh=0)),
                          ("+ tanh.approx swish (5e-4 rel)", dict(w=1, act=1, qk=1, v=1, tanh=5e-4))]:
            got = hoisted_scores(sd, tables, ds, cell, nbr, nbr_w, coords, opt)
            err = np.abs(sig(got + dp2) - ref)
            print("d=%3d %-34s max|err|=%.2e  mean=%.2e  (logit max err %.2e)" % (d, name, err.max(), err.mean(),
                  np.abs(got + dp2 - np.log(ref / (1 - ref))).max()))


if __name__ == "__main__":
    main()
