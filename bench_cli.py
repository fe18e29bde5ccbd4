#!/usr/bin/env python
"""bench_cli.py — wall clock of the index-building CLI on FASTA files (BASELINE configs 1 and 3 "on FASTA"):
   pgr-b200-make-frgdb <filelist> <prefix> --timing [--index-only] [--gpus N]
against the oracle doing FASTA -> canonical .mdb on all host cores, with the .mdb files compared byte for byte.
Writes the FASTA files under --dir (default /tmp/pgr_b200_cli).  One JSON line per run on stdout.
    python bench_cli.py --config 1            # one 1 Mb contig
    python bench_cli.py --config 3 --haps 94  # 94 x 50 Mb, one FASTA per haplotype (plain; --gz K compresses the first K)
"""
import argparse
import gzip
import json
import os
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import numpy as np  # noqa: E402

import bench_synth as S  # noqa: E402

CLI = os.path.join(ROOT, "pgr_tk_b200", "pgr-b200-make-frgdb")


def write_fasta(path, name, seq, gz=False, width=80):
    L = len(seq)
    nl = (L + width - 1) // width
    out = np.full(L + nl, 10, dtype=np.uint8)
    idx = np.arange(L, dtype=np.int64)
    out[idx + idx // width] = seq
    data = b">" + name.encode() + b"\n" + out.tobytes()
    if not data.endswith(b"\n"):
        data += b"\n"
    if gz:
        with gzip.open(path, "wb", compresslevel=1) as f:
            f.write(data)
    else:
        with open(path, "wb") as f:
            f.write(data)
    return len(data)


def host_cores():
    try:
        return len(os.sched_getaffinity(0))
    except AttributeError:
        return os.cpu_count() or 1


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--config", type=int, default=3)
    ap.add_argument("--haps", type=int, default=94)
    ap.add_argument("--hap-len", type=int, default=50_000_000)
    ap.add_argument("--gz", type=int, default=0, help="gzip the first K files")
    ap.add_argument("--gpus", default="1", help="comma list of GPU counts to run, e.g. 1,2")
    ap.add_argument("--readers", type=int, default=8)
    ap.add_argument("--dir", default="/tmp/pgr_b200_cli")
    ap.add_argument("--no-oracle", action="store_true")
    ap.add_argument("--full", action="store_true", help="also write the fragment store (.sdx/.frg)")
    args = ap.parse_args()
    os.makedirs(args.dir, exist_ok=True)
    t0 = time.perf_counter()
    paths, file_bytes, bases = [], 0, 0
    if args.config == 1:
        seq = S.rand_seq(np.random.default_rng(1), 1_000_000)
        p = os.path.join(args.dir, "c1.fa")
        file_bytes += write_fasta(p, "contig1", seq)
        paths.append(p)
        bases = len(seq)
    else:
        views, _, lens, owner = S.pangenome(args.hap_len, range(args.haps), threads=min(16, host_cores()))
        for h, v in enumerate(views):
            gz = h < args.gz
            p = os.path.join(args.dir, "hap%03d.fa" % h) + (".gz" if gz else "")
            file_bytes += write_fasta(p, "hap%03d" % h, v, gz=gz)
            paths.append(p)
        bases = int(sum(lens))
        del views, owner
    fl = os.path.join(args.dir, "files.txt")
    open(fl, "w").write("\n".join(paths) + "\n")
    gen_s = time.perf_counter() - t0
    base = {"config": args.config, "files": len(paths), "bases": bases, "file_bytes": file_bytes, "gz_files": args.gz, "synth_and_write_s": gen_s}
    ref_mdb = None
    for g in [int(x) for x in args.gpus.split(",")]:
        for mode in (["--index-only"], []) if args.full else (["--index-only"],):
            prefix = os.path.join(args.dir, "out_g%d%s" % (g, "_idx" if mode else "_full"))
            walls = []
            for rep in range(2):           # second run: page cache and CUDA context warm
                t0 = time.perf_counter()
                r = subprocess.run([CLI, fl, prefix, "--timing", "--readers", str(args.readers), "--gpus", str(g)] + mode, capture_output=True, text=True)
                walls.append(time.perf_counter() - t0)
                if r.returncode != 0:
                    print(json.dumps(dict(base, gpus=g, error=r.stderr[-500:])), flush=True)
                    break
            else:
                t = json.loads(r.stderr.strip().splitlines()[-1])
                rec = dict(base, gpus=g, mode="index-only" if mode else "full", process_wall_s=walls, cli=t,
                           gbases_per_s_process=bases / min(walls) / 1e9, gbases_per_s_in_process=bases / t["wall_s"] / 1e9)
                if ref_mdb is None:
                    ref_mdb = prefix + ".mdb"
                else:
                    rec["mdb_identical_to_first_run"] = open(prefix + ".mdb", "rb").read() == open(ref_mdb, "rb").read()
                print(json.dumps(rec), flush=True)
    if not args.no_oracle:
        import orc
        cores = host_cores()
        t0 = time.perf_counter()
        o = orc.Index(orc.mkspec(80, 56, 4, 64), 0)
        for p in paths:
            if p.endswith(".gz"):
                tmp = p[:-3] + ".tmp"
                open(tmp, "wb").write(gzip.open(p, "rb").read())
                o.load_fasta(tmp, nthreads=cores)
                os.remove(tmp)
            else:
                o.load_fasta(p, nthreads=cores)
        om = os.path.join(args.dir, "oracle.mdb")
        o.write_mdb(om)
        dt = time.perf_counter() - t0
        print(json.dumps(dict(base, impl="oracle (C++ port of the reference path: one file at a time, one sequence per thread, single-threaded inserts)", cores=cores,
                              wall_s=dt, gbases_per_s=bases / dt / 1e9,
                              mdb_identical_to_gpu=(open(om, "rb").read() == open(ref_mdb, "rb").read()) if ref_mdb else None)), flush=True)


if __name__ == "__main__":
    main()
