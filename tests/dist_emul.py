"""numpy emulation of distributed SpMVs driven by the product library's ghost plans (test helper).

For one rank: builds the vector segments in the library's local layout (F-local / nested order),
packs what the plan says every peer needs, exchanges, and multiplies with the local (diag, off-diag)
blocks.  The result must equal the corresponding rows of the serial product."""
import numpy as np

import pflare_b200


def nested_positions(lh, l_agg):
    """natural local index -> position in the rank's nested level vector (bottom level natural)."""
    NL = lh.no_levels
    LB = min(l_agg, NL)
    pos = {LB: np.arange(lh.levels[LB - 1]["n"])}
    for l in range(LB - 1, 0, -1):
        lv = lh.levels[l - 1]
        p = np.empty(lv["n"], dtype=np.int64)
        nf = lv["is_fine"].size
        p[lv["is_fine"]] = np.arange(nf)
        p[lv["is_coarse"]] = nf + pos[l + 1]
        pos[l] = p
    return pos


def serial_products(H, l, xs):
    lv = H.levels[l - 1]
    return {"AFF": lv.A_ff @ xs["f"], "AFC": lv.A_fc @ xs["c"], "R": lv.R @ xs["v"], "P": lv.P @ xs["c"]}


def distributed_products(H, parts, rank, world, dev, l_agg, exchange):
    """Returns the worst relative error over levels / operators of this rank's rows."""
    lh = parts[rank]
    pos = nested_positions(lh, l_agg)
    worst = 0.0
    rng = np.random.default_rng(5)
    for l in range(1, min(l_agg, H.no_levels)):
        lvs = H.levels[l - 1]
        nfg, ncg, ng = lvs.is_fine.size, lvs.is_coarse.size, lvs.n
        xs = {"f": rng.random(nfg), "c": rng.random(ncg), "v": rng.random(ng)}
        ref = serial_products(H, l, xs)
        lv = lh.levels[l - 1]
        rV, rF = lh.rangesV[l - 1], lh.rangesF[l - 1]
        rC = lh.rangesV[l]
        nf = lv["is_fine"].size
        # my vector segments in the library's layout
        seg_f = xs["f"][rF[rank]:rF[rank + 1]]                            # F-local
        xc_nat = xs["c"][rC[rank]:rC[rank + 1]]
        seg_c = np.empty_like(xc_nat)
        seg_c[pos[l + 1]] = xc_nat                                        # nested order of level l+1
        v_nat = xs["v"][rV[rank]:rV[rank + 1]]
        seg_bf = v_nat[lv["is_fine"]]                                     # b_f, F-local
        for which, seg, key, rows in ((pflare_b200.AFF, seg_f, "AFF", (rF, None)),
                                      (pflare_b200.AFC, seg_c, "AFC", (rF, None)),
                                      (pflare_b200.P, seg_c, "P", (rV, "W")),
                                      (pflare_b200.R, seg_bf, "R", (rC, "Z"))):
            plan = dev.ghost_plan(l, which)
            sendbufs, off = [], 0
            for p in range(world):
                cnt = int(plan["send_count"][p])
                sendbufs.append(seg[plan["send_idx"][off:off + cnt]])
                off += cnt
            got = exchange(sendbufs, plan["recv_count"])
            op = lv["ops"][which]
            xg = np.zeros(op.garray.size)
            for p in range(world):
                cnt = int(plan["recv_count"][p])
                if cnt:
                    xg[plan["recv_off"][p]:plan["recv_off"][p] + cnt] = got[p]
            # local product in natural local numbering of the operator's column space
            if key == "AFF":
                y = op.diag @ seg_f + (op.offdiag @ xg if op.offdiag is not None else 0)
            elif key in ("AFC", "P"):
                y = op.diag @ xc_nat + (op.offdiag @ xg if op.offdiag is not None else 0)
            else:   # R = [Z I]: ghosts are F points of other ranks; local columns natural
                y = op.diag @ v_nat + (op.offdiag @ xg if op.offdiag is not None else 0)
            r0 = rows[0]
            want = ref[key][r0[rank]:r0[rank + 1]]
            den = np.linalg.norm(ref[key]) or 1.0
            worst = max(worst, float(np.linalg.norm(y - want) / den))
    return worst
