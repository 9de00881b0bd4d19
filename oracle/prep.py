"""Per-cell neighbourhood preparation for imputation -- restates higashi/Impute.py:16-26
(moving_avg) and :78-121 (prep_one), and the kNN cell graph of
higashi/Higashi_wrapper.py:1303-1324 (get_cell_neighbor) + :1520-1526 (weight_dict).

TEST INFRASTRUCTURE (see oracle/__init__.py).
"""
import numpy as np
from scipy.sparse import csr_matrix, vstack

PDF0 = 0.3989422804014327            # scipy.stats.norm.pdf(0) = 1/sqrt(2 pi)


def norm_pdf(x):
    """scipy.stats.norm.pdf: exp(-x^2/2)/sqrt(2 pi)."""
    return float(np.exp(-0.5 * x * x) / np.sqrt(2.0 * np.pi))


def moving_avg(adj, moving_range):
    """Impute.moving_avg (Impute.py:16-26): Gaussian row-shift smoothing; range 0 => adj * pdf(0).

    Rows shifted past the first/last row replicate that row."""
    adj_origin = adj.copy()
    adj = adj.copy() * np.float64(PDF0)
    for i in range(moving_range * 2):
        before = vstack([adj_origin[0, :]] * (i + 1) + [adj_origin[:-(i + 1), :]])
        after = vstack([adj_origin[i + 1:, :]] + [adj_origin[-1, :]] * (i + 1))
        adj = adj + (after + before) * np.float64(norm_pdf((i + 1) / moving_range))
    return adj


def l1_normalize_rows(adj):
    """sklearn.preprocessing.normalize(adj, norm='l1', axis=1): divide each row by sum|x|; zero rows stay."""
    adj = csr_matrix(adj, copy=True)
    s = np.asarray(abs(adj).sum(axis=1)).reshape(-1)
    s[s == 0] = 1.0
    adj.data = adj.data / np.repeat(s, np.diff(adj.indptr))
    return adj


def prep_one(matrices, local_transfer_range, nbr_cells=None, nbr_weights=None):
    """Impute.prep_one for ONE chromosome of ONE cell (Impute.py:78-121).

    matrices: callable cell_id(1-based) -> scipy CSR fp32 of that chromosome.
    nbr_cells / nbr_weights: the cell's neighbour list (1-based ids, self first) and the
    weights weight_dict[(nbr, cell)]; None => unweighted (mtx = A_cell)."""
    if nbr_cells is None:
        raise ValueError("pass nbr_cells=[cell] for the unweighted case")
    if nbr_weights is None:
        mtx = matrices(nbr_cells[0])
    else:
        mtx = 0
        for nb, w in zip(nbr_cells, nbr_weights):
            mtx = mtx + w * matrices(int(nb))          # sequential fp32 CSR sums (Impute.py:85-88)
    adj = moving_avg(csr_matrix(mtx), local_transfer_range)
    adj.data = np.log1p(adj.data)
    adj = l1_normalize_rows(adj).astype("float32")
    return adj


def cell_neighbors(cell_embeddings, neighbor_num):
    """Higashi.get_cell_neighbor(start=0) (Higashi_wrapper.py:1303-1324).

    neighbor_num = config['neighbor_num'] + 1 (wrapper:530).  Returns (nbr int64 [n,neighbor_num]
    1-based ids, self first; w float32 [n,neighbor_num]).

    sklearn.metrics.pairwise_distances(metric='euclidean') on float32 input: squared norms and
    the cross term are accumulated in float64 (`_euclidean_distances_upcast`), clipped at 0,
    the diagonal zeroed, sqrt taken, and the result returned as float32."""
    v = np.asarray(cell_embeddings)
    x = v.astype(np.float64)
    xx = (x * x).sum(-1)
    d2 = xx[:, None] + xx[None, :] - 2.0 * (x @ x.T)
    np.maximum(d2, 0, out=d2)
    np.fill_diagonal(d2, 0.0)
    distance = np.sqrt(d2).astype(v.dtype if v.dtype == np.float32 else np.float64)
    distance_sorted = np.sort(distance, axis=-1)
    distance = distance / np.quantile(distance_sorted[:, 1:neighbor_num].reshape(-1), q=0.25)
    nbr = np.argsort(distance, axis=-1, kind="stable")[:, :neighbor_num]
    w = np.take_along_axis(distance, nbr, axis=-1)
    w = np.exp(-w)
    w = w / w.sum(-1, keepdims=True)
    return nbr.astype(np.int64) + 1, w
