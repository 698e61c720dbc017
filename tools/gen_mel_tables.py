"""Generates lighting-asr_b200/csrc/mel_static_default.inc: the mel projection of LASR's default
option set (16 kHz, 512-point FFT, 80 bins, 20 Hz .. Nyquist; lasr/data/datatrans.py:45-70) as
straight-line macro code, one section per warp of the fused kernel's phase B.

The weights are computed with the same float32 torch expressions as torchaudio's get_mel_banks
(TA:436-511), so they are bit-identical to what the reference multiplies with; the 1/4 that
undoes the un-normalised real-FFT split (|2X|^2) is folded in (exact, power of two).
Run:  python tools/gen_mel_tables.py   (needs torch; output is committed)."""
import importlib
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def tables(num_mel=80, padded=512, sample_freq=16000.0, low=20.0, high=0.0):
    fe = importlib.import_module("lighting-asr_b200.frontend")
    W = fe._torch_mel_banks(num_mel, padded, sample_freq, low, high).numpy()      # [num_mel, padded/2]
    nb = padded // 2
    seg = np.full(nb, -1, dtype=np.int64)
    up = np.zeros(nb, dtype=np.float32)
    dn = np.zeros(nb, dtype=np.float32)
    import math
    nyq = 0.5 * sample_freq
    hi = high + nyq if high <= 0 else high
    mel_lo = 1127.0 * math.log(1.0 + low / 700.0)
    mel_hi = 1127.0 * math.log(1.0 + hi / 700.0)
    delta = (mel_hi - mel_lo) / (num_mel + 1)
    bin_w = sample_freq / padded
    for k in range(nb):
        m = 1127.0 * math.log(1.0 + bin_w * k / 700.0)
        if m <= mel_lo:
            s = -1
        else:  # same rule as b200fe_plan_create: number of centres strictly below mel(k)
            s = 0
            while s <= num_mel and mel_lo + (s + 1) * delta < m:
                s += 1
        nz = set(np.nonzero(W[:, k])[0].tolist())
        assert nz <= {s, s - 1}, (k, s, nz)
        seg[k] = min(s, num_mel + 1)
        if 0 <= s < num_mel:
            up[k] = W[s, k]
        if 1 <= s <= num_mel:
            dn[k] = W[s - 1, k]
    assert np.all(np.diff(seg[seg >= 0]) >= 0)
    seg_start = np.zeros(num_mel + 3, dtype=np.int64)
    k = 0
    for s in range(num_mel + 2):
        while k < nb and seg[k] < s:
            k += 1
        seg_start[s] = k
    seg_start[num_mel + 2] = nb
    return W, seg, up, dn, seg_start


def group_cost(jb, je, seg_start, up, dn):
    """Instructions of group_code(jb, je): one multiply(-add) per non-zero weight, one store per bin, one add where a bin uses two
    chains, one float4 load per four FFT bins of the group's range."""
    n = 0
    for j in range(jb, je):
        t = sum(1 for k in range(int(seg_start[j]), int(seg_start[j + 1])) if up[k] != 0.0)
        t += sum(1 for k in range(int(seg_start[j + 1]), int(seg_start[j + 2])) if dn[k] != 0.0)
        n += t + 1 + (1 if t >= 4 else 0)
    return n + ((int(seg_start[je + 1]) - 1) >> 2) - (int(seg_start[jb]) >> 2) + 1


def groups(seg_start, num_mel, n_groups=8, up=None, dn=None):
    """Contiguous bin groups (one per warp of phase B) that minimise the LARGEST group: the phase ends at a CTA barrier, so the
    slowest warp sets its length.  Exact dynamic programme over the cut points."""
    if up is None:                     # the warp-specialised kernel keeps the older segment-length model
        cost = np.array([(seg_start[j + 2] - seg_start[j]) * 0.5 + 6.0 for j in range(num_mel)])
        tot = cost.sum()
        g = [0]
        j, acc = 0, 0.0
        for w in range(1, n_groups):
            target = tot * w / n_groups
            while j < num_mel and acc + cost[j] * 0.5 < target:
                acc += cost[j]
                j += 1
            g.append(j)
        g.append(num_mel)
        return g
    INF = 10 ** 9
    c = [[group_cost(a, b, seg_start, up, dn) if b > a else INF for b in range(num_mel + 1)] for a in range(num_mel + 1)]
    best = [[INF] * (num_mel + 1) for _ in range(n_groups + 1)]
    arg = [[0] * (num_mel + 1) for _ in range(n_groups + 1)]
    best[0][0] = 0
    for w in range(1, n_groups + 1):
        for b in range(w, num_mel + 1):
            for a in range(w - 1, b):
                v = max(best[w - 1][a], c[a][b])
                if v < best[w][b]:
                    best[w][b], arg[w][b] = v, a
    g, b = [num_mel], num_mel
    for w in range(n_groups, 0, -1):
        b = arg[w][b]
        g.append(b)
    return g[::-1]


def group_code(w, jb, je, seg_start, up, dn):
    """Straight-line code of one bin group: every float4 of the power row the group touches is loaded first (the loads are
    independent of everything else, so they overlap instead of each sitting in front of its first use), then every bin is one or
    two multiply-add chains over its non-zero weights -- no zero-initialised accumulators (an `0.f + x` survives optimisation: it
    is not an identity for x = -0) and no adds of empty partial sums."""
    comp = "xyzw"
    k_lo, k_hi = int(seg_start[jb]), int(seg_start[je + 1])
    lines = ["MGROUP_BEGIN(%d)" % w]
    if k_hi > k_lo:
        lines.append("const float4 " + ", ".join("q%d = pcol[%d]" % (c, c) for c in range(k_lo >> 2, ((k_hi - 1) >> 2) + 1)) + ";")
    for j in range(jb, je):
        terms = [(up[k] * 0.25, k) for k in range(int(seg_start[j]), int(seg_start[j + 1])) if up[k] != 0.0]
        terms += [(dn[k] * 0.25, k) for k in range(int(seg_start[j + 1]), int(seg_start[j + 2])) if dn[k] != 0.0]
        p = lambda k: "q%d.%s" % (k >> 2, comp[k & 3])
        if not terms:
            lines.append("emit_bin(orow, %d, 0.f);" % j)
            continue
        chains = [terms] if len(terms) < 4 else [terms[0::2], terms[1::2]]
        body = []
        for ci, ch in enumerate(chains):
            body.append("float r%d = %.9ef * %s;" % (ci, ch[0][0], p(ch[0][1])))
            for wt, k in ch[1:]:
                body.append("r%d = fmaf(%.9ef, %s, r%d);" % (ci, wt, p(k), ci))
        lines.append("{ " + " ".join(body) + " emit_bin(orow, %d, %s); }" % (j, "r0" if len(chains) == 1 else "r0 + r1"))
    lines.append("MGROUP_END(%d)" % w)
    return lines


def main():
    num_mel = 80
    W, seg, up, dn, seg_start = tables(num_mel)
    out = []
    out.append("// GENERATED by tools/gen_mel_tables.py -- do not edit.")
    out.append("// Mel projection for sample_frequency=16000, padded window 512, num_mel_bins=80, low_freq=20, high_freq=0.")
    out.append("// Device sections: one straight-line function per bin group (float4 loads of the power row first, then one or two")
    out.append("// multiply-add chains per bin); one section per CTA shape (B200FE_WARPS warps per CTA = number of bin groups).")
    out.append("// Warp-specialised kernel: MK(k, up, down) accumulates FFT bin k, MEND0() closes the leading segment of a group,")
    out.append("// MEND(j): bin j is complete (its down-slope segment just ended).")
    out.append("#define B200FE_STATIC_NMEL %d" % num_mel)
    out.append("#ifdef B200FE_MEL_HOST_TABLES")
    out.append("static const short kStaticSegStart[%d] = {%s};" % (len(seg_start), ", ".join(str(int(v)) for v in seg_start)))
    out.append("static const float kStaticUp[256] = {%s};" % ", ".join("%.9ef" % (v * 0.25) for v in up))
    out.append("static const float kStaticDn[256] = {%s};" % ", ".join("%.9ef" % (v * 0.25) for v in dn))
    out.append("#endif")
    for ng in (8, 6):
        grp = groups(seg_start, num_mel, ng, up, dn)
        out.append("#if B200FE_WARPS == 8 || B200FE_WARPS == 4" if ng == 8 else "#if B200FE_WARPS == %d" % ng)   # 16-frame tiles: one group per half-warp
        out.append("#ifdef B200FE_MEL_HOST_TABLES")
        out.append("static const short kStaticGrpBegin[9] = {%s};" % ", ".join(str(v) for v in (grp + [num_mel] * 9)[:9]))
        out.append("#endif")
        out.append("#ifdef B200FE_MEL_DEVICE_CODE")
        for w in range(ng):
            out.extend(group_code(w, grp[w], grp[w + 1], seg_start, up, dn))
        out.append("#endif")
        out.append("#endif")
    # warp-specialised kernel (fbank_ws_kernel.cuh): one group per epilogue warp
    out.append("#ifdef B200FE_MEL_WS_CODE")
    for ng in (4, 3):
        grp = groups(seg_start, num_mel, ng)
        out.append("#if B200FE_WS_EPI == %d" % ng)
        out.append("#define B200FE_WS_GRP_BEGIN {%s}" % ", ".join(str(v) for v in (grp + [num_mel] * 5)[:5]))
        for w in range(ng):
            jb, je = grp[w], grp[w + 1]
            out.append("MGROUP_BEGIN(%d, %d, %d)" % (w, jb, je))
            for s_ in range(jb, je + 1):
                for k in range(int(seg_start[s_]), int(seg_start[s_ + 1])):
                    u = up[k] * 0.25 if s_ < je else 0.0
                    d = dn[k] * 0.25 if s_ > jb else 0.0
                    out.append("MK(%d, %.9ef, %.9ef)" % (k, u, d))
                out.append("MEND0()" if s_ == jb else "MEND(%d)" % (s_ - 1))
            out.append("MGROUP_END(%d)" % w)
        out.append("#endif")
    out.append("#endif")
    grp = groups(seg_start, num_mel, 8, up, dn)
    path = os.path.join(ROOT, "lighting-asr_b200", "csrc", "mel_static_default.inc")
    open(path, "w").write("\n".join(out) + "\n")
    nk = sum(int(seg_start[grp[w + 1] + 1] - seg_start[grp[w]]) for w in range(8))
    print("wrote", path, "groups", grp, "k reads per frame", nk, "nnz", int((W > 0).sum()))


if __name__ == "__main__":
    main()
