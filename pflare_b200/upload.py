"""Upload hook: walk a Hierarchy exactly as the reference-side uploader would walk ``air_multigrid_data``
(SURVEY.md section 8b "Upload hook": after ``PCSetUp(pcmg)``, /root/reference/src/AIR_MG_Setup.F90:1178)
and hand every operator to a sink exposing set_level / set_csr / set_diag / set_poly / finalize.
Both the CPU oracle and the CUDA C-ABI binding implement that sink, so they are fed identical
operators."""

AFF, AFC, ACF, ACC, INV_AFF, INV_ACC, R, P, COARSE = range(9)


def _feed_inverse(sink, our_level, which, inv):
    if inv.kind == "csr":
        sink.set_csr(our_level, which, inv.mat)
    elif inv.kind == "diag":
        sink.set_diag(our_level, which, inv.diag)
    elif inv.kind == "poly":
        sink.set_poly(our_level, which, inv.inverse_type, inv.coeffs, inv.diag_scale)
    else:
        raise ValueError(inv.kind)


def feed(H, sink):
    NL = H.no_levels
    full = bool(getattr(getattr(H, "options", None), "full_smoothing_up_and_down", False))
    if full:
        # -pc_air_full_smoothing_up_and_down: the smoother is inv_A_ff(level) applied to coarse_matrix(level) on all
        # unknowns (src/AIR_MG_Setup.F90:1014-1074): the hook hands over coarse_matrix, inv_A_ff, R and P per level
        sink.set_option("full_smoothing_up_and_down", 1)
    coarse_its = int(getattr(getattr(H, "options", None), "mg_coarse_ksp_max_it", 1) or 1)
    if coarse_its > 1:
        sink.set_option("mg_coarse_ksp_max_it", coarse_its)
    for l, lv in enumerate(H.levels):
        ol = l + 1
        sink.set_level(ol, lv.n, lv.is_fine, lv.is_coarse, lv.smooth_order)
        if full:
            sink.set_csr(ol, COARSE, lv.A)
            _feed_inverse(sink, ol, INV_AFF, lv.inv_A_ff)
            sink.set_csr(ol, R, lv.R)
            sink.set_csr(ol, P, lv.P)
            continue
        sink.set_csr(ol, AFF, lv.A_ff)
        sink.set_csr(ol, AFC, lv.A_fc)
        if lv.A_cf is not None and lv.A_cc is not None:
            sink.set_csr(ol, ACF, lv.A_cf)
            sink.set_csr(ol, ACC, lv.A_cc)
            _feed_inverse(sink, ol, INV_ACC, lv.inv_A_cc)
        _feed_inverse(sink, ol, INV_AFF, lv.inv_A_ff)
        sink.set_csr(ol, R, lv.R)
        sink.set_csr(ol, P, lv.P)
    n = H.coarse_matrix.shape[0]
    sink.set_level(NL, n, [], [], [])
    sink.set_csr(NL, COARSE, H.coarse_matrix)
    _feed_inverse(sink, NL, INV_AFF, H.inv_coarse)
    sink.finalize()
    return sink
