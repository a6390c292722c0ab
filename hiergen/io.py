"""Hierarchy container I/O: one .npz per hierarchy (fixtures under tests/golden/, rank-to-rank
hand-over through /dev/shm in bench.py).  Keys are flat: ``L{l}_{name}_{indptr|indices|data|shape}``.
"""
import numpy as np
import scipy.sparse as sp

from .setup import AirOptions, Hierarchy, Inverse, Level


def _put_csr(d, key, m):
    m = m.tocsr()
    d[key + "_indptr"] = m.indptr.astype(np.int32)
    d[key + "_indices"] = m.indices.astype(np.int32)
    d[key + "_data"] = m.data.astype(np.float64)
    d[key + "_shape"] = np.asarray(m.shape, dtype=np.int64)


def _get_csr(d, key):
    shape = tuple(int(v) for v in d[key + "_shape"])
    m = sp.csr_matrix((d[key + "_data"], d[key + "_indices"], d[key + "_indptr"]), shape=shape)
    m.has_sorted_indices = True
    return m


def _put_inv(d, key, inv):
    d[key + "_kind"] = np.asarray({"csr": 1, "diag": 2, "poly": 3}[inv.kind])
    d[key + "_type"] = np.asarray(int(inv.inverse_type))
    d[key + "_dscale"] = np.asarray(int(bool(inv.diag_scale)))
    if inv.kind == "csr":
        _put_csr(d, key + "_mat", inv.mat)
    elif inv.kind == "diag":
        d[key + "_diag"] = np.asarray(inv.diag, dtype=np.float64)
    else:
        d[key + "_coeffs"] = np.asarray(inv.coeffs, dtype=np.float64)


def _get_inv(d, key):
    kind = {1: "csr", 2: "diag", 3: "poly"}[int(d[key + "_kind"])]
    inv = Inverse(kind, inverse_type=int(d[key + "_type"]), diag_scale=bool(int(d[key + "_dscale"])))
    if kind == "csr":
        inv.mat = _get_csr(d, key + "_mat")
    elif kind == "diag":
        inv.diag = np.array(d[key + "_diag"])
    else:
        inv.coeffs = np.array(d[key + "_coeffs"])
    return inv


def to_dict(H, with_A=True):
    d = {"no_levels": np.asarray(H.no_levels), "full_smoothing": np.asarray(int(bool(H.options.full_smoothing_up_and_down)))}
    if with_A:
        _put_csr(d, "A", H.A)
    for l, lv in enumerate(H.levels):
        k = "L%d" % (l + 1)
        d[k + "_n"] = np.asarray(lv.n)
        d[k + "_is_fine"] = np.asarray(lv.is_fine, dtype=np.int32)
        d[k + "_is_coarse"] = np.asarray(lv.is_coarse, dtype=np.int32)
        d[k + "_smooth"] = np.asarray(lv.smooth_order, dtype=np.int32)
        d[k + "_affdiag"] = np.asarray(int(lv.aff_diag))
        for name in ("A_ff", "A_fc", "R", "P"):
            _put_csr(d, k + "_" + name, getattr(lv, name))
        _put_inv(d, k + "_inv_A_ff", lv.inv_A_ff)
        if lv.A is not None:
            _put_csr(d, k + "_A", lv.A)
        if lv.A_cf is not None and lv.A_cc is not None:
            _put_csr(d, k + "_A_cf", lv.A_cf)
            _put_csr(d, k + "_A_cc", lv.A_cc)
            _put_inv(d, k + "_inv_A_cc", lv.inv_A_cc)
    _put_csr(d, "coarse_matrix", H.coarse_matrix)
    _put_inv(d, "inv_coarse", H.inv_coarse)
    return d


def from_dict(d):
    NL = int(d["no_levels"])
    levels = []
    for l in range(1, NL):
        k = "L%d" % l
        lv = Level(n=int(d[k + "_n"]), is_fine=np.array(d[k + "_is_fine"]), is_coarse=np.array(d[k + "_is_coarse"]),
                   A_ff=_get_csr(d, k + "_A_ff"), A_fc=_get_csr(d, k + "_A_fc"), inv_A_ff=_get_inv(d, k + "_inv_A_ff"),
                   R=_get_csr(d, k + "_R"), P=_get_csr(d, k + "_P"), smooth_order=[int(v) for v in d[k + "_smooth"]],
                   aff_diag=bool(int(d[k + "_affdiag"])))
        if (k + "_A_indptr") in d:
            lv.A = _get_csr(d, k + "_A")
        if (k + "_A_cc_indptr") in d:
            lv.A_cf = _get_csr(d, k + "_A_cf")
            lv.A_cc = _get_csr(d, k + "_A_cc")
            lv.inv_A_cc = _get_inv(d, k + "_inv_A_cc")
        levels.append(lv)
    cm = _get_csr(d, "coarse_matrix")
    A = _get_csr(d, "A") if "A_indptr" in d else (None if levels else cm)
    full = bool(int(d["full_smoothing"])) if "full_smoothing" in d else False
    return Hierarchy(A=A, levels=levels, coarse_matrix=cm, inv_coarse=_get_inv(d, "inv_coarse"),
                     options=AirOptions(full_smoothing_up_and_down=full))


def save(path, H, compressed=True, **extra):
    d = to_dict(H)
    d.update(extra)
    (np.savez_compressed if compressed else np.savez)(path, **d)


def load(path):
    with np.load(path) as z:
        d = {k: z[k] for k in z.files}
    return from_dict(d), d
