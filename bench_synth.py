"""Seeded synthetic inputs of the BASELINE.json configs, shared by bench.py and bench_configs.py.

config 3 pangenome: one ancestor (uniform ACGT, seed 7); every haplotype derived independently from it with SNPs at rate
1e-3, indels (length geometric, mean 3) at 1e-4, 20 inversions and 20 tandem duplications of 1-50 kb (seed 700+h).
Positions are drawn directly (binomial count, uniform positions) instead of thresholding one uniform per base, and the
indels / duplications are applied in ONE pass over the sequence, so that a 50 Mb haplotype takes ~0.2 s.

"assembly-like" decoration (what real assemblies add to uniform ACGT): runs of N (gaps), soft-masked lower case,
microsatellites ((AT)n and friends, reverse-complement palindromes >= k) and inverted repeats.
"""
import concurrent.futures as cf

import numpy as np

ACGT = np.frombuffer(b"ACGT", dtype=np.uint8)
COMP = np.arange(256, dtype=np.uint8)
for _a, _b in zip(b"ACGTacgt", b"TGCAtgca"):
    COMP[_a] = _b


def rand_seq(rng, L):
    return ACGT[rng.integers(0, 4, size=L, dtype=np.uint8)]


def ancestor(L, seed=7):
    return rand_seq(np.random.default_rng(seed), L)


def haplotype_pieces(anc, seed, snp=1e-3, indel=1e-4, n_sv=20):
    """-> list of arrays whose concatenation is the haplotype"""
    rng = np.random.default_rng(seed)
    L = len(anc)
    s = anc.copy()
    k = int(rng.binomial(L, snp))
    pos = rng.integers(0, L, size=k)
    s[pos] = ACGT[rng.integers(0, 4, size=k, dtype=np.uint8)]
    for _ in range(n_sv):                                  # inversions, in place
        ln = int(rng.integers(1000, 50000))
        a = int(rng.integers(0, max(1, L - ln)))
        s[a:a + ln] = COMP[s[a:a + ln]][::-1]
    # one left-to-right pass: indels and tandem duplications as edit events sorted by position
    m = int(rng.binomial(L, indel))
    ipos = rng.integers(1, L, size=m)
    ilen = rng.geometric(1.0 / 3.0, size=m)
    idel = rng.random(m) < 0.5
    dlen = rng.integers(1000, 50000, size=n_sv)
    dpos = np.array([int(rng.integers(0, max(1, L - int(d)))) for d in dlen], dtype=np.int64)
    ev = [(int(p), 0 if d else 1, int(ln)) for p, ln, d in zip(ipos, ilen, idel)] + [(int(p) + int(d), 2, int(d)) for p, d in zip(dpos, dlen)]
    ev.sort()
    pieces, last = [], 0
    for p, kind, ln in ev:
        if p < last:
            continue
        pieces.append(s[last:p])
        if kind == 0:                                      # deletion
            last = min(L, p + ln)
        elif kind == 1:                                    # insertion of random bases
            pieces.append(rand_seq(rng, ln))
            last = p
        else:                                              # tandem duplication of the ln bases before p
            pieces.append(s[p - ln:p])
            last = p
    pieces.append(s[last:])
    return pieces


def pangenome(L, hap_ids, alloc=None, threads=8, anc=None):
    """haplotypes `hap_ids` of the config-3 pangenome.  alloc(nbytes) -> (uint8 array, base address) supplies the memory the
    sequences are written into (e.g. a pinned HostBuffer); default numpy.  Returns (views, ptrs, lens, buffer_owner)."""
    if anc is None:
        anc = ancestor(L)
    hap_ids = list(hap_ids)

    def gen(h):
        return haplotype_pieces(anc, 700 + h)

    with cf.ThreadPoolExecutor(max_workers=max(1, threads)) as ex:
        allp = list(ex.map(gen, hap_ids))
    lens = [int(sum(len(p) for p in ps)) for ps in allp]
    offs, off = [], 0
    for ln in lens:
        offs.append(off)
        off += (ln + 63) & ~63
    if alloc is None:
        buf = np.zeros(max(off, 1), dtype=np.uint8)
        base = buf.ctypes.data
        owner = buf
    else:
        owner = alloc(max(off, 1))
        buf, base = owner.array, owner.ptr

    def put(i):
        np.concatenate(allp[i], out=buf[offs[i]:offs[i] + lens[i]])
        allp[i] = None

    with cf.ThreadPoolExecutor(max_workers=max(1, threads)) as ex:
        list(ex.map(put, range(len(hap_ids))))
    views = [buf[o:o + ln] for o, ln in zip(offs, lens)]
    ptrs = [base + o for o in offs]
    return views, ptrs, lens, owner


def decorate_assembly_like(seq, seed, n_frac=0.02, lower_frac=0.3, repeat_every=50_000):
    """in place: ~n_frac of the bases in N runs of 10 kb .. 5 Mb (log-uniform), ~lower_frac soft-masked in 200 b .. 20 kb runs,
    and every ~repeat_every bases one low-complexity insert: (AT)n / (CG)n / (ACGT)n / an inverted repeat (stem 60-400)
    — all of them contain reverse-complement palindromes of length >= 56, the case the minimizer machine special-cases
    (shmmrutils.rs:477)."""
    rng = np.random.default_rng(seed)
    L = len(seq)
    # low-complexity inserts first (N runs and case are laid over them)
    n_rep = max(1, L // repeat_every)
    for p in rng.integers(0, max(1, L - 2000), size=n_rep):
        p = int(p)
        kind = int(rng.integers(0, 4))
        if kind == 3:
            stem = int(rng.integers(60, 400))
            a = seq[p:p + stem].copy()
            seq[p + stem:p + 2 * stem] = COMP[a][::-1][:max(0, min(stem, L - p - stem))]
        else:
            unit = (b"AT", b"CG", b"ACGT")[kind]
            ln = int(rng.integers(60, 1200))
            ln = min(ln, L - p)
            rep = np.frombuffer(unit * (ln // len(unit) + 1), dtype=np.uint8)[:ln]
            seq[p:p + ln] = rep
    # soft masking
    covered = 0
    while covered < lower_frac * L:
        ln = int(rng.integers(200, 20000))
        a = int(rng.integers(0, max(1, L - ln)))
        seq[a:a + ln] |= 0x20
        covered += ln
    # gaps
    covered = 0
    while covered < n_frac * L:
        ln = int(min(max(10_000, 1.5 * n_frac * L), 10 ** rng.uniform(4.0, 6.7)))
        a = int(rng.integers(0, max(1, L - ln)))
        seq[a:a + ln] = ord("N")
        covered += ln
    return seq
