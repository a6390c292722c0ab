"""PETSc-binary hierarchy container: the file a PFLARE build dumps after ``PCSetUp`` (with the C helper of
``shim/pflare_b200_petsc.c``: ``MatView`` / ``ISView`` / ``VecView`` on one ``PetscViewerBinaryOpen`` viewer) and that this
package reads back and walks through the upload hook -- so a hierarchy built by the reference's own setup (gfortran +
PETSc, not available in this image) can be fed to the CUDA path and to the oracle, together with a ``PCApply`` input /
output pair that pins vector-level parity against the real reference.

File = PETSc binary objects back to back (big endian; class ids and layouts as PETSc writes them, verified against the
reference's own fixtures ``tests/data/*`` -- SURVEY.md section 8c):

    Mat (AIJ)  int32 [1211216, M, N, nnz], M row lengths, nnz int32 columns, nnz float64 values
    Vec        int32 [1211214, n], n float64
    IS         int32 [1211218, n], n int32

Object order (names follow ``air_multigrid_data``, /root/reference/src/AIR_Data_Type.F90:284-360):

    IS  header      [0x50464231, no_levels, full_smoothing_up_and_down, has_apply_pair]
    for our_level = 1 .. no_levels - 1:
        IS  meta    [n, n_fine, n_coarse, has_c_smoothing]
        IS  IS_fine_index ; IS IS_coarse_index ; IS smooth_order_levels
        full smoothing:  Mat coarse_matrix(our_level)
        otherwise:       Mat A_ff ; Mat A_fc ; [Mat A_cf ; Mat A_cc ; INV inv_A_cc]
        INV inv_A_ff ; Mat restrictors ; Mat prolongators
    Mat coarse_matrix(no_levels) ; INV inv_A_ff(no_levels)
    [Vec b ; Vec x]   one PCApply input / output pair produced by the reference (the vector-level pin)

    INV := IS [kind, PCPFLAREINVType, diag_scale, ncoef]  followed by
           kind 1 (assembled AIJ): Mat | kind 2 (MATDIAGONAL): Vec | kind 3 (MatShell polynomial): Vec real, Vec imag

Index lists are local to the rank (global - rstart) like the upload hook's; a parallel run dumps the global matrices
(``MatView`` of an MPIAIJ Mat writes the whole matrix), which ``hiergen.partition`` splits again.
"""
import numpy as np
import scipy.sparse as sp

MAT_ID, VEC_ID, IS_ID = 1211216, 1211214, 1211218
MAGIC = 0x50464231


class Inverse:
    def __init__(self, kind, mat=None, diag=None, inverse_type=1, coeffs=None, diag_scale=False):
        self.kind, self.mat, self.diag, self.inverse_type, self.coeffs, self.diag_scale = kind, mat, diag, inverse_type, coeffs, diag_scale


class Level:
    def __init__(self):
        self.n = 0
        self.is_fine = self.is_coarse = None
        self.smooth_order = []
        self.A_ff = self.A_fc = self.A_cf = self.A_cc = self.A = self.R = self.P = None
        self.inv_A_ff = self.inv_A_cc = None


class Options:
    def __init__(self, full):
        self.full_smoothing_up_and_down = bool(full)


class Hierarchy:
    """What ``pflare_b200.upload.feed`` walks (same attribute names as hiergen's Hierarchy)."""

    def __init__(self):
        self.A = None
        self.levels = []
        self.coarse_matrix = None
        self.inv_coarse = None
        self.options = Options(False)
        self.b = self.x = None       # optional PCApply pair

    @property
    def no_levels(self):
        return len(self.levels) + 1

    def sizes(self):
        return [lv.n for lv in self.levels] + [self.coarse_matrix.shape[0]]


# ---------------------------------------------------------------------------------------------- writer
class Writer:
    def __init__(self, f):
        self.f = f

    def is_(self, v):
        v = np.asarray(v, dtype=np.int64).ravel()
        np.array([IS_ID, v.size], dtype=">i4").tofile(self.f)
        v.astype(">i4").tofile(self.f)

    def vec(self, v):
        v = np.asarray(v, dtype=np.float64).ravel()
        np.array([VEC_ID, v.size], dtype=">i4").tofile(self.f)
        v.astype(">f8").tofile(self.f)

    def mat(self, m):
        m = sp.csr_matrix(m)
        m.sort_indices()
        np.array([MAT_ID, m.shape[0], m.shape[1], m.nnz], dtype=">i4").tofile(self.f)
        np.diff(m.indptr).astype(">i4").tofile(self.f)
        m.indices.astype(">i4").tofile(self.f)
        m.data.astype(">f8").tofile(self.f)

    def inv(self, inv):
        kind = {"csr": 1, "diag": 2, "poly": 3}[inv.kind]
        co = np.atleast_2d(np.asarray(inv.coeffs, dtype=np.float64)) if inv.kind == "poly" else np.zeros((0, 1))
        if inv.kind == "poly" and co.shape[0] == 1 and np.asarray(inv.coeffs).ndim == 1:
            co = co.T
        self.is_([kind, int(inv.inverse_type), int(bool(inv.diag_scale)), co.shape[0]])
        if kind == 1:
            self.mat(inv.mat)
        elif kind == 2:
            self.vec(inv.diag)
        else:
            self.vec(co[:, 0])
            self.vec(co[:, 1] if co.shape[1] > 1 else np.zeros(co.shape[0]))


def save_hierarchy(path, H, b=None, x=None):
    full = bool(getattr(H.options, "full_smoothing_up_and_down", False))
    with open(path, "wb") as f:
        w = Writer(f)
        w.is_([MAGIC, H.no_levels, int(full), int(b is not None and x is not None)])
        for lv in H.levels:
            has_c = lv.A_cf is not None and lv.A_cc is not None and not full
            w.is_([lv.n, len(lv.is_fine), len(lv.is_coarse), int(has_c)])
            w.is_(lv.is_fine); w.is_(lv.is_coarse); w.is_(lv.smooth_order)
            if full:
                w.mat(lv.A)
            else:
                w.mat(lv.A_ff); w.mat(lv.A_fc)
                if has_c:
                    w.mat(lv.A_cf); w.mat(lv.A_cc); w.inv(lv.inv_A_cc)
            w.inv(lv.inv_A_ff)
            w.mat(lv.R); w.mat(lv.P)
        w.mat(H.coarse_matrix)
        w.inv(H.inv_coarse)
        if b is not None and x is not None:
            w.vec(b); w.vec(x)


# ---------------------------------------------------------------------------------------------- reader
class Reader:
    def __init__(self, buf):
        self.buf, self.off = buf, 0

    def _i4(self, n):
        v = np.frombuffer(self.buf, ">i4", n, self.off).astype(np.int64)
        self.off += 4 * n
        return v

    def _expect(self, cid, what):
        got = int(self._i4(1)[0])
        if got != cid:
            raise ValueError("PETSc binary: expected %s (class id %d) at offset %d, found %d" % (what, cid, self.off - 4, got))

    def is_(self):
        self._expect(IS_ID, "IS")
        n = int(self._i4(1)[0])
        return self._i4(n).astype(np.int32)

    def vec(self):
        self._expect(VEC_ID, "Vec")
        n = int(self._i4(1)[0])
        v = np.frombuffer(self.buf, ">f8", n, self.off).astype(np.float64)
        self.off += 8 * n
        return v

    def mat(self):
        self._expect(MAT_ID, "Mat")
        M, N, nnz = [int(v) for v in self._i4(3)]
        rl = self._i4(M)
        cols = self._i4(nnz).astype(np.int32)
        vals = np.frombuffer(self.buf, ">f8", nnz, self.off).astype(np.float64)
        self.off += 8 * nnz
        indptr = np.concatenate(([0], np.cumsum(rl))).astype(np.int32)
        m = sp.csr_matrix((vals, cols, indptr), shape=(M, N))
        m.sort_indices()
        return m

    def inv(self):
        kind, itype, dscale, ncoef = [int(v) for v in self.is_()]
        if kind == 1:
            return Inverse("csr", mat=self.mat(), inverse_type=itype, diag_scale=bool(dscale))
        if kind == 2:
            return Inverse("diag", diag=self.vec(), inverse_type=itype, diag_scale=bool(dscale))
        if kind == 3:
            re, im = self.vec(), self.vec()
            if re.size != ncoef or im.size != ncoef:
                raise ValueError("PETSc binary: polynomial inverse with %d coefficients, header says %d" % (re.size, ncoef))
            newton = itype in (2, 3)
            return Inverse("poly", inverse_type=itype, coeffs=np.stack((re, im), axis=1) if newton else re.reshape(-1, 1),
                           diag_scale=bool(dscale))
        raise ValueError("PETSc binary: unknown inverse kind %d" % kind)


def load_hierarchy(path):
    r = Reader(open(path, "rb").read())
    hdr = r.is_()
    if hdr.size != 4 or int(hdr[0]) != MAGIC:
        raise ValueError("%s is not a pflare_b200 hierarchy dump" % path)
    NL, full, has_pair = int(hdr[1]), bool(hdr[2]), bool(hdr[3])
    H = Hierarchy()
    H.options = Options(full)
    for _ in range(NL - 1):
        lv = Level()
        n, nf, nc, has_c = [int(v) for v in r.is_()]
        lv.n = n
        lv.is_fine, lv.is_coarse = r.is_(), r.is_()
        lv.smooth_order = [int(v) for v in r.is_()]
        if lv.is_fine.size != nf or lv.is_coarse.size != nc:
            raise ValueError("PETSc binary: index set sizes disagree with the level header")
        if full:
            lv.A = r.mat()
        else:
            lv.A_ff, lv.A_fc = r.mat(), r.mat()
            if has_c:
                lv.A_cf, lv.A_cc = r.mat(), r.mat()
                lv.inv_A_cc = r.inv()
        lv.inv_A_ff = r.inv()
        lv.R, lv.P = r.mat(), r.mat()
        H.levels.append(lv)
    H.coarse_matrix = r.mat()
    H.inv_coarse = r.inv()
    if has_pair:
        H.b, H.x = r.vec(), r.vec()
    if r.off != len(r.buf):
        raise ValueError("PETSc binary: %d trailing bytes" % (len(r.buf) - r.off))
    return H
