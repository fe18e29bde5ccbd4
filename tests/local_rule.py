"""numpy model of the LOCAL (stateless) formulation of sequence_to_shmmrs that the CUDA kernels implement.

The reference (shmmrutils.rs:417-556) is a sequential state machine.  The kernels use an equivalent local rule
(DESIGN.md "Local rule"); this file is that rule written with numpy so the equivalence can be property-tested
against the oracle on CPU (tests/test_local_rule.py).  It is a test helper, not a product path.
"""
import numpy as np

U64 = np.uint64
_LUT = np.full(256, 4, dtype=np.uint8)
_LUT[0:4] = [0, 1, 2, 3]
for ch, v in zip(b"ACGTacgt", [0, 1, 2, 3, 0, 1, 2, 3]):
    _LUT[ch] = v


def u64hash(key):
    key = key.astype(U64)
    with np.errstate(over="ignore"):
        key = (~key) + (key << U64(21))
        key = key ^ (key >> U64(24))
        key = (key + (key << U64(3))) + (key << U64(8))
        key = key ^ (key >> U64(14))
        key = (key + (key << U64(2))) + (key << U64(4))
        key = key ^ (key >> U64(28))
        key = key + (key << U64(31))
    return key


def kmer_keys(seq, k):
    """For every position p: (hash, strand, palindrome) of the k-mer made of the last k VALID bases at or
    before p (zero fill when fewer), exactly as the rolling registers of shmmrutils.rs:446-476 hold them."""
    s = np.frombuffer(bytes(seq), dtype=np.uint8)
    L = len(s)
    code = _LUT[s]
    valid = code < 4
    vcodes = code[valid].astype(U64)
    nv = len(vcodes)
    # registers after j valid bases (j = 1..nv): f = last k codes, newest at LSB; r = complement, newest at bit k-1
    f0 = np.zeros(nv + 1, dtype=U64)
    f1 = np.zeros(nv + 1, dtype=U64)
    r0 = np.zeros(nv + 1, dtype=U64)
    r1 = np.zeros(nv + 1, dtype=U64)
    pad = np.concatenate([np.zeros(k, dtype=U64), vcodes])
    cpad = np.concatenate([np.zeros(k, dtype=U64), U64(3) - vcodes])  # zero fill is 0 in the r registers too
    for i in range(k):  # bit i of f = code of the base i steps back; bit (k-1-i) of r
        b = pad[k - i: k - i + nv]
        c = cpad[k - i: k - i + nv]
        f0[1:] |= (b & U64(1)) << U64(i)
        f1[1:] |= ((b >> U64(1)) & U64(1)) << U64(i)
        r0[1:] |= (c & U64(1)) << U64(k - 1 - i)
        r1[1:] |= ((c >> U64(1)) & U64(1)) << U64(k - 1 - i)
    vcount = np.cumsum(valid)  # number of valid bases in [0..p]
    F0, F1, R0, R1 = f0[vcount], f1[vcount], r0[vcount], r1[vcount]
    pal = (F0 == R0) & (F1 == R1)
    fwd = ~(R0 < F0)
    a = np.where(fwd, F0, R0)
    b = np.where(fwd, F1, R1)
    h = u64hash(a) ^ u64hash(b ^ U64(0xAD12CF59))
    return h, (~fwd).astype(np.uint8), pal


def window_select(x, w):
    """tie-inclusive window-minimum selection: i selected iff x[i] is a minimum of some full window of w
    consecutive elements of x.  Returns a boolean mask."""
    n = len(x)
    sel = np.zeros(n, dtype=bool)
    if n < w:
        return sel
    win = np.lib.stride_tricks.sliding_window_view(x, w)  # (n-w+1, w)
    m = win.min(axis=1)
    # M[i] = max over windows containing i
    M = np.zeros(n, dtype=x.dtype)
    for off in range(w):
        seg = M[off: off + len(m)]
        np.maximum(seg, m, out=seg)
    return M == x


def level0(seq, w, k, rid=0):
    """level-0 minimizers (position-ordered list of (x, y)); returns (xs, ys, had_palindrome)."""
    L = len(seq)
    h, strand, pal = kmer_keys(seq, k)
    with np.errstate(over="ignore"):
        x = (h << U64(8)) | U64(k)
    had_pal = bool(pal[k:].any()) if L > k else False
    pos = np.arange(L, dtype=np.int64)
    y = (U64(rid) << U64(32)) | (pos.astype(U64) << U64(1)) | strand.astype(U64)
    if L <= k:
        return x[:0], y[:0], had_pal
    E = L - w + k  # rule (2) is active for pos < E (shmmrutils.rs:516-519)
    lo = k
    hi = min(L, E)  # local region: windows inside [lo, hi)
    sel_pos = []
    if hi - lo >= w:
        m = window_select(x[lo:hi], w)
        sel_pos = list(np.nonzero(m)[0] + lo)
    # tail replay (only when w > k): rescans only
    q = sel_pos[-1] if sel_pos else None
    for p in range(max(hi, lo), L):
        fire = (p == k + w - 1) if q is None else (p == q + w)
        if fire:
            a = p - w + 1
            seg = x[a: p + 1]
            mn = seg.min()
            for j in np.nonzero(seg == mn)[0]:
                sel_pos.append(a + int(j))
            q = sel_pos[-1]
    sp = np.array(sel_pos, dtype=np.int64)
    return x[sp], y[sp], had_pal


MAXU = U64(0xFFFFFFFFFFFFFFFF)


def reduce_local(xs, ys, r, padding):
    if padding:
        padx = np.full(r - 1, MAXU, dtype=U64)
        xs = np.concatenate([padx, xs, padx])
        ys = np.concatenate([padx, ys, padx])
    m = window_select(xs, r)
    return xs[m], ys[m]


def span_filter(xs, ys, min_span):
    n = len(xs)
    if n <= 2:
        return xs, ys
    pos = ((ys & U64(0xFFFFFFFF)) >> U64(1)).astype(np.int64)
    keep = np.ones(n, dtype=bool)
    d_prev = (pos[1:-1] - pos[:-2]) & 0xFFFFFFFF
    d_next = (pos[2:] - pos[1:-1]) & 0xFFFFFFFF
    keep[1:-1] = (d_prev > min_span) & (d_next > min_span) & (xs[:-2] != xs[1:-1]) & (xs[1:-1] != xs[2:])
    return xs[keep], ys[keep]


def sequence_to_shmmrs_local(rid, seq, w, k, r, min_span, padding=False):
    """returns (xs, ys, had_palindrome); only guaranteed equal to the reference when had_palindrome is False"""
    xs, ys, had_pal = level0(seq, w, k, rid)
    if r > 1:
        xs, ys = reduce_local(xs, ys, r, padding)
        xs, ys = reduce_local(xs, ys, r, padding)
    xs, ys = span_filter(xs, ys, min_span)
    return xs, ys, had_pal


def sketch_local(rid, seq, k, r, min_span):
    L = len(seq)
    h, strand, pal = kmer_keys(seq, k)
    thr = (0xFFFFFFFFFFFFFFFF >> 4) >> r
    pos = np.arange(L, dtype=np.int64)
    keep = (pos >= k) & (~pal) & (h < U64(thr))
    with np.errstate(over="ignore"):
        x = (h << U64(8)) | U64(k)
    y = (U64(rid) << U64(32)) | (pos.astype(U64) << U64(1)) | strand.astype(U64)
    return span_filter(x[keep], y[keep], min_span)
