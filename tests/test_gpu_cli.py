"""GPU: the C++ host mirror / CLI (pgr-b200-make-frgdb, same command line as pgr-make-frgdb) writes the reference's
file formats: <prefix>.mdb equals the oracle's canonical .mdb byte for byte and equals the reference's committed fixture
as a map; <prefix>.midx equals the fixture's."""
import os
import subprocess

import pytest

import orc
import pgr_tk_b200 as pg

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, "tests", "golden")
CLI = os.path.join(ROOT, "pgr_tk_b200", "pgr-b200-make-frgdb")


def test_make_frgdb_cli_matches_fixture_and_oracle(tmp_path):
    if not os.path.exists(CLI):
        subprocess.check_call(["make", "-C", os.path.join(ROOT, "pgr_tk_b200", "host")])
    fa = os.path.join(GOLDEN, "test_seqs.fa")
    fl = tmp_path / "files.txt"
    fl.write_text(fa + "\n")
    prefix = str(tmp_path / "out")
    subprocess.check_call([CLI, str(fl), prefix], cwd=ROOT)
    _, ref_map, _ = orc.read_mdb_py(os.path.join(GOLDEN, "test_seqs_frag.mdb"))
    spec, got_map, order = orc.read_mdb_py(prefix + ".mdb")
    assert spec == (80, 56, 4, 64, 0) and got_map == ref_map and order == sorted(order)
    o = orc.Index(orc.mkspec(), 0)
    o.load_fasta(fa)
    o.write_mdb(str(tmp_path / "o.mdb"))
    assert open(prefix + ".mdb", "rb").read() == open(tmp_path / "o.mdb", "rb").read()
    ref_midx = [l.rstrip("\n").split("\t") for l in open(os.path.join(GOLDEN, "test_seqs_frag.midx"))]
    got_midx = [l.rstrip("\n").split("\t") for l in open(prefix + ".midx")]
    assert len(got_midx) == len(ref_midx) == 66
    for g, r in zip(got_midx, ref_midx):
        assert g[:3] == r[:3] and os.path.basename(g[3]) == os.path.basename(r[3])
    # the fragment store: decoded content equals the reference's own .sdx/.frg (the deflate streams themselves differ
    # between zlib and the reference's miniz; the inflated bincode payloads are byte-identical)
    import frag_format as ff
    rcs, raddr, rseqs = ff.read_sdx(os.path.join(GOLDEN, "test_seqs_frag.sdx"))
    rpay = ff.read_frg_chunks(os.path.join(GOLDEN, "test_seqs_frag.frg"), raddr)
    gcs, gaddr, gseqs = ff.read_sdx(prefix + ".sdx")
    gpay = ff.read_frg_chunks(prefix + ".frg", gaddr)
    assert gcs == rcs == 256 and gpay == rpay
    assert [a[2] for a in gaddr] == [a[2] for a in raddr]
    assert [a[0] for a in gaddr] == [sum(x[1] for x in gaddr[:i]) for i in range(len(gaddr))]
    strip = lambda q: [(x["name"], x["id"], x["seq_frag_range"], x["len"], os.path.basename(x["source"])) for x in q]
    assert strip(gseqs) == strip(rseqs)
    frags = ff.decode_chunks(gpay)
    recs = orc.parse_fasta(fa)
    assert all(ff.get_seq(frags, 56, x) == recs[i][1] for i, x in enumerate(gseqs))


def test_make_frgdb_cli_append_and_flags(tmp_path):
    """two files (load_from_fastx then append_from_fastx, ext.rs:152-199): sids and fragment ids continue; non-default spec"""
    if not os.path.exists(CLI):
        subprocess.check_call(["make", "-C", os.path.join(ROOT, "pgr_tk_b200", "host")])
    fa1, fa2 = os.path.join(GOLDEN, "test_seqs.fa"), os.path.join(GOLDEN, "test_rev.fa")
    fl = tmp_path / "files.txt"
    fl.write_text(fa1 + "\n" + fa2 + "\n")
    prefix = str(tmp_path / "out2")
    subprocess.check_call([CLI, str(fl), prefix, "-w", "48", "-k", "56", "-r", "4", "--min-span", "12"], cwd=ROOT)
    o = orc.Index(orc.mkspec(48, 56, 4, 12), 0)
    o.load_fasta(fa1)
    o.load_fasta(fa2)
    o.write_mdb(str(tmp_path / "o.mdb"))
    assert open(prefix + ".mdb", "rb").read() == open(tmp_path / "o.mdb", "rb").read()
    assert len(open(prefix + ".midx").readlines()) == 68
    # the .mdb can be loaded back by the library
    assert pg.ShmmrIndex.read_mdb(prefix + ".mdb").as_map() == o.as_map()
    # the fragment store of the two files: fragment ids continue across files; content equals the oracle's, and every
    # sequence is reconstructed from it
    import frag_format as ff
    from test_frag_format import oracle_db
    recs = orc.parse_fasta(fa1) + orc.parse_fasta(fa2)
    exp = oracle_db(recs, (48, 56, 4, 12))
    cs, addr, seqs = ff.read_sdx(prefix + ".sdx")
    frags = ff.decode_chunks(ff.read_frg_chunks(prefix + ".frg", addr))
    assert frags == exp.frags
    assert [(x["name"], x["id"], x["seq_frag_range"], x["len"]) for x in seqs] == [(x["name"], x["id"], x["seq_frag_range"], x["len"]) for x in exp.seqs]
    assert [os.path.basename(x["source"]) for x in seqs] == ["test_seqs.fa"] * 66 + ["test_rev.fa"] * 2
    assert all(ff.get_seq(frags, 56, x) == recs[i][1] for i, x in enumerate(seqs))


# ---- pgr-b200-query (pgr-query.rs) ---------------------------------------------------------------------------------------
import sys  # noqa: E402

import numpy as np  # noqa: E402

sys.path.insert(0, os.path.join(ROOT, "oracle"))
import query_post_oracle as qpo  # noqa: E402

QCLI = os.path.join(ROOT, "pgr_tk_b200", "pgr-b200-query")
ACGT = np.frombuffer(b"ACGT", dtype=np.uint8)


def _write_fasta(path, recs):
    with open(path, "w") as f:
        for name, seq in recs:
            f.write(">%s some description\n" % name)
            s = seq.decode()
            for i in range(0, len(s), 80):
                f.write(s[i:i + 80] + "\n")


def _oracle_targets(o, q, penalty, **kw):
    osid, otco, osc, ocho, ohits = o.query_fragment_to_hps(q, penalty, **kw)
    targets = []
    for t, sid in enumerate(osid):
        alns = []
        for c in range(int(otco[t]), int(otco[t + 1])):
            hp = [((int(h["qb"]), int(h["qe"]), int(h["qo"])), (int(h["tb"]), int(h["te"]), int(h["to"]))) for h in ohits[int(ocho[c]):int(ocho[c + 1])]]
            alns.append((float(osc[c]), hp))
        targets.append((int(sid), alns))
    return targets


def test_query_cli_matches_oracle_postprocessing(tmp_path):
    if not os.path.exists(QCLI):
        subprocess.check_call(["make", "-C", os.path.join(ROOT, "pgr_tk_b200", "host")])
    rng = np.random.default_rng(101)
    anc = ACGT[rng.integers(0, 4, size=160000)]
    comp = np.zeros(256, dtype=np.uint8)
    comp[[65, 67, 71, 84]] = [84, 71, 67, 65]
    haps = []
    for h in range(6):
        s = anc.copy()
        m = np.nonzero(rng.random(len(s)) < 2e-3)[0]
        s[m] = ACGT[rng.integers(0, 4, size=len(m))]
        if h % 2:     # an inversion and a large deletion: several chains / orientations per target
            a = int(rng.integers(40000, 60000))
            s = np.concatenate([s[:a], comp[s[a:a + 15000]][::-1], s[a + 15000:a + 40000], s[a + 70000:]])
        haps.append(s.tobytes())
    db_fa = str(tmp_path / "db.fa")
    _write_fasta(db_fa, [("hap%d" % i, s) for i, s in enumerate(haps)])
    queries = [("q_fwd", haps[0][30000:130000]), ("q_rev", bytes(comp[np.frombuffer(haps[2][20000:90000], dtype=np.uint8)][::-1])),
               ("q_none", ACGT[rng.integers(0, 4, size=5000)].tobytes())]
    q_fa = str(tmp_path / "q.fa")
    _write_fasta(q_fa, queries)
    o = orc.Index(orc.mkspec(), 0)
    o.add_batch(list(range(len(haps))), haps)
    seq_info = {i: ("hap%d" % i, db_fa) for i in range(len(haps))}
    kw = dict(max_count=128, max_count_query=128, max_count_target=128, max_aln_span=8)
    for tol, bed in ((100000, False), (1000, True)):
        prefix = str(tmp_path / ("out_%d" % tol))
        cmd = [QCLI, db_fa, q_fa, prefix, "--fastx-file", "--merge-range-tol", str(tol)] + (["--bed-summary"] if bed else [])
        subprocess.check_call(cmd, cwd=ROOT)
        n_regions = 0
        for idx, (qn, qs) in enumerate(queries):
            merged = qpo.merge_query_hits(_oracle_targets(o, qs, 0.025, **kw), tol)
            text, subs = qpo.hit_lines(idx, qn, len(qs), merged, seq_info, bed=bed)
            got = open("%s.%03d.%s" % (prefix, idx, "hit.bed" if bed else "hit")).read()
            assert got == text, (tol, idx)
            fa_exp = "".join(">%s\n%s\n" % (nm, (qpo.reverse_complement(haps[sid][b:e]) if ori else haps[sid][b:e]).decode()) for nm, sid, b, e, ori in subs)
            assert open("%s.%03d.fa" % (prefix, idx)).read() == fa_exp
            n_regions += len(subs)
        assert n_regions >= 8
    # index files + --only-summary (the .mdb/.midx of pgr-b200-make-frgdb read back)
    fl = tmp_path / "files.txt"
    fl.write_text(db_fa + "\n")
    subprocess.check_call([CLI, str(fl), str(tmp_path / "idx")], cwd=ROOT)
    prefix = str(tmp_path / "out_idx")
    subprocess.check_call([QCLI, str(tmp_path / "idx"), q_fa, prefix, "--only-summary"], cwd=ROOT)
    for idx in range(len(queries)):
        assert open("%s.%03d.hit" % (prefix, idx)).read() == open("%s.%03d.hit" % (str(tmp_path / "out_100000"), idx)).read()
        assert not os.path.exists("%s.%03d.fa" % (prefix, idx))
    # the same with the .mdb left on disk (read_mdb_file_to_frag_locations + query_fragment_to_hps_from_mmap_file, ext.rs:87-150,285-342)
    prefix_m = str(tmp_path / "out_mmap")
    subprocess.check_call([QCLI, str(tmp_path / "idx"), q_fa, prefix_m, "--only-summary", "--mdb-resident"], cwd=ROOT)
    for idx in range(len(queries)):
        assert open("%s.%03d.hit" % (prefix_m, idx)).read() == open("%s.%03d.hit" % (prefix, idx)).read()


# ---- pgr-b200-pbundle-decomp (pgr-pbundle-decomp.rs) ----------------------------------------------------------------------
import bundles_oracle as bo  # noqa: E402
import pbundle_oracle as pbo  # noqa: E402

PCLI = os.path.join(ROOT, "pgr_tk_b200", "pgr-b200-pbundle-decomp")


def test_pbundle_decomp_cli_matches_oracle(tmp_path):
    if not os.path.exists(PCLI):
        subprocess.check_call(["make", "-C", os.path.join(ROOT, "pgr_tk_b200", "host")])
    from test_gpu_bundles import repetitive_locus
    haps = repetitive_locus(29, 10, n_units=5, unit_len=7000, flank=25000)
    fa = str(tmp_path / "locus.fa")
    _write_fasta(fa, [("hap_%02d" % (len(haps) - i), s) for i, s in enumerate(haps)])     # names sort differently from sids
    spec_t = (48, 56, 4, 12)
    o = orc.Index(orc.mkspec(*spec_t), 0)
    o.add_batch(list(range(len(haps))), haps)
    keys, offs, _ = o.export()
    kk = keys.reshape(-1, 2)
    cnt = {(int(kk[i, 0]), int(kk[i, 1])): int(offs[i + 1] - offs[i]) for i in range(len(kk))}
    adj = o.adj_list(0)
    adj_t = [(int(r["sid"]), (int(r["a0"]), int(r["a1"]), int(r["ori0"])), (int(r["b0"]), int(r["b1"]), int(r["ori1"]))) for r in adj]
    smps = {}
    for sid, s in enumerate(haps):
        pairs, _, _ = o.raw_query(s)
        smps[sid] = [(int(p["h0"]), int(p["h1"]), int(p["bgn"]), int(p["end"]), int(p["ori"])) for p in pairs]
    seq_info = [(sid, len(s), "hap_%02d" % (len(haps) - sid)) for sid, s in enumerate(haps)]
    for extra, (cut, merge, branch) in (([], (2500, 10000, 8)), (["--bundle-length-cutoff", "500", "--bundle-merge-distance", "2000", "--min-branch-size", "3"], (500, 2000, 3))):
        pb, _ = bo.get_principal_bundles_from_adj_list(cnt, adj_t, branch)
        pbid, vmap = pbo.principal_bundles_with_id(pb, smps)
        prefix = str(tmp_path / ("pb%d" % cut))
        cmd = [PCLI, fa, prefix] + extra
        subprocess.check_call(cmd, cwd=ROOT)
        bed, summ = pbo.decomposition_files(" ".join(cmd), seq_info, smps, pbid, vmap, spec_t[1], cut, merge)
        assert open(prefix + ".bed").read() == bed
        assert open(prefix + ".ctg.summary.tsv").read() == summ
        assert bed.count("\n") > 20 and ":R" in bed and ":U" in bed
        # MAP-graph files: the whole graph, its index, the principal graph with bundle tags
        fmap = o.as_map()
        assert open(prefix + ".mapg.gfa").read() == pbo.gfa_text(adj_t, fmap, spec_t[1])
        assert open(prefix + ".mapg.idx").read() == pbo.mapg_idx_text(spec_t + (0,), [(sid, ln, ctg, fa) for sid, ln, ctg in seq_info], fmap)
        _, flt = bo.get_principal_bundles_from_adj_list(cnt, adj_t, branch)
        plain = {(v[0], v[1]): (b, v[2], p) for b, path in enumerate(pb) for p, v in enumerate(path)}
        pm = open(prefix + ".pmapg.gfa").read()
        assert pm == pbo.gfa_text(flt, fmap, spec_t[1], plain) and "\tBN:i:" in pm
        # the .pdb: same content as the oracle's bundles; and -p reads it back to the same decomposition
        head, d_pbid, d_vmap = pbo.decode_pdb(open(prefix + ".pdb", "rb").read())
        assert head == spec_t + (branch, 0) and d_pbid == [(b, o_, list(v)) for b, o_, v in pbid] and d_vmap == vmap
        assert open(prefix + ".pdb", "rb").read() == pbo.encode_pdb(*spec_t, branch, 0, pbid, vmap)
        prefix2 = prefix + "_from_pdb"
        subprocess.check_call([PCLI, fa, prefix2, "-p", prefix + ".pdb", "-w", "80", "--bundle-length-cutoff", str(cut), "--bundle-merge-distance", str(merge)], cwd=ROOT)
        assert open(prefix2 + ".bed").read().split("\n")[1:] == bed.split("\n")[1:]
        assert open(prefix2 + ".ctg.summary.tsv").read() == summ
        assert not os.path.exists(prefix2 + ".mapg.gfa") and not os.path.exists(prefix2 + ".pdb")
        # --include: only the listed contigs are decomposed (against the bundles of the whole file)
        inc = tmp_path / ("inc%d.txt" % cut)
        chosen = ["hap_03", "hap_09", "hap_01"]
        inc.write_text("\n".join(chosen) + "\n")
        prefix3 = prefix + "_inc"
        subprocess.check_call([PCLI, fa, prefix3, "-i", str(inc)] + extra, cwd=ROOT)
        keep = lambda text: [l for l in text.split("\n")[1:] if l and l.split("\t")[0] in chosen]
        assert keep(open(prefix3 + ".bed").read()) == keep(bed) and len(keep(bed)) > 3
        assert [l for l in open(prefix3 + ".bed").read().split("\n")[1:] if l] == keep(bed)
        assert keep(open(prefix3 + ".ctg.summary.tsv").read()) == keep(summ) and len(keep(summ)) == 3


def test_query_cli_frg_backend_and_reference_fragment_store(tmp_path):
    """--frg-file: target sub-sequences come out of the .sdx/.frg store.  (1) a store written by pgr-b200-make-frgdb gives
    the same files as the FASTX back end; (2) the REFERENCE's own fixture store (test_seqs_frag.*) is read back correctly."""
    import shutil
    for cli in (CLI, QCLI):
        if not os.path.exists(cli):
            subprocess.check_call(["make", "-C", os.path.join(ROOT, "pgr_tk_b200", "host")])
    rng = np.random.default_rng(202)
    anc = ACGT[rng.integers(0, 4, size=90000)]
    haps = []
    for h in range(4):
        s = anc.copy()
        m = np.nonzero(rng.random(len(s)) < 3e-3)[0]
        s[m] = ACGT[rng.integers(0, 4, size=len(m))]
        haps.append(s.tobytes())
    db_fa = str(tmp_path / "db.fa")
    _write_fasta(db_fa, [("h%d" % i, s) for i, s in enumerate(haps)])
    q_fa = str(tmp_path / "q.fa")
    _write_fasta(q_fa, [("qa", haps[1][10000:70000]), ("qb", haps[3][40000:80000])])
    fl = tmp_path / "files.txt"
    fl.write_text(db_fa + "\n")
    subprocess.check_call([CLI, str(fl), str(tmp_path / "store")], cwd=ROOT)
    subprocess.check_call([QCLI, db_fa, q_fa, str(tmp_path / "fx"), "--fastx-file"], cwd=ROOT)
    subprocess.check_call([QCLI, str(tmp_path / "store"), q_fa, str(tmp_path / "fr"), "--frg-file"], cwd=ROOT)
    for idx in range(2):
        for ext in ("hit", "fa"):
            a, b = open("%s.%03d.%s" % (tmp_path / "fx", idx, ext)).read(), open("%s.%03d.%s" % (tmp_path / "fr", idx, ext)).read()
            assert a == b and len(a) > 100
    # (2) the reference's fixture
    for ext in ("mdb", "midx", "sdx", "frg"):
        shutil.copy(os.path.join(GOLDEN, "test_seqs_frag." + ext), str(tmp_path / ("ref." + ext)))
    recs = orc.parse_fasta(os.path.join(GOLDEN, "test_seqs.fa"))
    by_name = {n: s for n, s in recs}
    q2 = str(tmp_path / "q2.fa")
    _write_fasta(q2, [("q0", recs[0][1]), ("q1", qpo.reverse_complement(recs[5][1][200:3000]))])
    subprocess.check_call([QCLI, str(tmp_path / "ref"), q2, str(tmp_path / "rf"), "--frg-file"], cwd=ROOT)
    n_checked = 0
    for idx in range(2):
        rows = [l.rstrip("\n").split("\t") for l in open("%s.%03d.hit" % (tmp_path / "rf", idx)) if not l.startswith("#")]
        fa = open("%s.%03d.fa" % (tmp_path / "rf", idx)).read().split("\n")
        assert len(rows) >= 10 and len(fa) == 2 * len(rows) + 1
        for r, (hdr, seq) in zip(rows, zip(fa[0::2], fa[1::2])):
            ctg, b, e, ori = r[7], int(r[8]), int(r[9]), int(r[10])
            sub = by_name[ctg][b:e]
            assert hdr == ">" + r[11] and seq.encode() == (qpo.reverse_complement(sub) if ori else sub)
            n_checked += 1
    assert n_checked >= 20


def test_make_frgdb_cli_sharded_pipeline_same_files(tmp_path):
    """--devices 0,0,0 (the multi-GPU build with three shards on one device), several readers, plain and .gz inputs, a FASTQ:
    every output file equals the single-GPU run's byte for byte (.frg: same inflated payloads), --index-only writes the same
    .mdb/.midx and no store, --timing prints one JSON line"""
    import gzip
    import json
    import numpy as np
    if not os.path.exists(CLI):
        subprocess.check_call(["make", "-C", os.path.join(ROOT, "pgr_tk_b200", "host")])
    rng = np.random.default_rng(77)
    acgt = np.frombuffer(b"ACGT", dtype=np.uint8)
    anc = acgt[rng.integers(0, 4, size=150_000)]
    paths = []
    for i in range(7):
        s = anc.copy()
        m = rng.random(len(s)) < 0.004
        s[m] = acgt[rng.integers(0, 4, size=int(m.sum()))]
        recs = b""
        for j in range(int(rng.integers(1, 4))):
            a = int(rng.integers(0, 40_000))
            part = s[a:a + int(rng.integers(30_000, 100_000))].tobytes()
            recs += b">h%d_c%d sample\n" % (i, j) + b"\n".join(part[q:q + 70] for q in range(0, len(part), 70)) + b"\n"
        p = str(tmp_path / ("h%d.fa" % i)) + (".gz" if i % 2 else "")
        if p.endswith(".gz"):
            with gzip.open(p, "wb") as f:
                f.write(recs)
        else:
            open(p, "wb").write(recs)
        paths.append(p)
    fq = str(tmp_path / "reads.fq")
    open(fq, "wb").write(b"".join(b"@r%d\n%s\n+\n%s\n" % (i, anc[i * 9000:i * 9000 + 8000].tobytes(), b"I" * 8000) for i in range(4)))
    paths.insert(3, fq)
    fl = tmp_path / "files.txt"
    fl.write_text("\n".join(paths) + "\n")
    one, sh, io = str(tmp_path / "one"), str(tmp_path / "sharded"), str(tmp_path / "indexonly")
    subprocess.check_call([CLI, str(fl), one, "--readers", "1"], cwd=ROOT)
    r = subprocess.run([CLI, str(fl), sh, "--devices", "0,0,0", "--readers", "4", "--timing"], cwd=ROOT, capture_output=True, text=True, check=True)
    t = json.loads(r.stderr.strip().splitlines()[-1])
    assert t["files"] == 8 and t["gpus"] == 3 and t["bases"] > 0 and t["wall_s"] > 0
    subprocess.check_call([CLI, str(fl), io, "--index-only"], cwd=ROOT)
    for ext in (".mdb", ".midx", ".sdx"):
        assert open(one + ext, "rb").read() == open(sh + ext, "rb").read(), ext
    assert open(one + ".frg", "rb").read() == open(sh + ".frg", "rb").read()
    assert open(io + ".mdb", "rb").read() == open(one + ".mdb", "rb").read() and open(io + ".midx", "rb").read() == open(one + ".midx", "rb").read()
    assert not os.path.exists(io + ".frg") and not os.path.exists(io + ".sdx")
    # against the oracle: same records (the reference's FASTQ reader does not yield the last record of the file), same map
    sys_path = os.path.join(ROOT, "oracle")
    import sys
    sys.path.insert(0, sys_path)
    import fastx_oracle as fo
    recs = []
    for p in paths:
        buf = gzip.open(p, "rb").read() if p.endswith(".gz") else open(p, "rb").read()
        recs += fo.parse_fastx(buf)
    assert len(recs) == len(open(one + ".midx").read().splitlines())
    o = orc.Index(orc.mkspec(), 0)
    o.add_batch(list(range(len(recs))), [s for _, s in recs])
    o.write_mdb(str(tmp_path / "o.mdb"))
    assert open(one + ".mdb", "rb").read() == open(tmp_path / "o.mdb", "rb").read()
