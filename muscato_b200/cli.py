"""Stage-compatible host driver for the hot path (Python mirror of steps 5-12 of
cmd/muscato/main.go:1029-1051): same config.json, same files.

  python -m muscato_b200.cli <LogDir/config.json> [--device N] [--from-fastq]

Reads  TempDir/reads_sorted.txt.sz (or, with --from-fastq, builds it from ReadFileName like
       prepReads does), Config.GeneFileName, Config.GeneIdFileName
Writes TempDir/matches.txt.sz, Config.ResultsFileName and the non-match fastq
       (cmd/muscato_nonmatch/main.go:66-71 naming).
"""
from __future__ import annotations

import argparse
import os
import sys

from . import formats, sz
from .config import Config
from .engine import HotPath


def run(config_path: str, device: int = 0, from_fastq: bool = False) -> dict:
    cfg = Config.from_json(config_path).apply_defaults()
    if not cfg.TempDir:
        raise SystemExit("TempDir must be set in the config")
    os.makedirs(cfg.TempDir, exist_ok=True)
    rs_path = os.path.join(cfg.TempDir, "reads_sorted.txt.sz")
    if from_fastq:
        with open(cfg.ReadFileName, "rb") as f:
            seqs, counts, names = formats.prep_reads_uniqify(f.read(), cfg.MinReadLength, cfg.MaxReadLength)
        sz.write_file(rs_path, b"".join(s + b"\t" + c + b"\t" + n + b"\n" for s, c, n in zip(seqs, counts, names)))
    else:
        if not os.path.exists(rs_path) and os.path.exists(rs_path[:-3]):
            rs_path = rs_path[:-3]
        seqs, counts, names = formats.load_reads_sorted(rs_path)
    targets = formats.load_targets(cfg.GeneFileName)
    gnames, glens = formats.load_gene_ids(cfg.GeneIdFileName)
    with HotPath(cfg, device=device) as hp:
        hp.set_reads(seqs)
        hp.set_targets(targets)
        hp.run()
        m = hp.fetch()
        st = hp.stats()
    lines = formats.matches_lines(m, seqs, targets)
    sz.write_file(os.path.join(cfg.TempDir, "matches.txt.sz"), b"".join(ln + b"\n" for ln in lines))
    res = formats.results_lines(m, seqs, counts, names, targets, gnames, glens)
    results = cfg.ResultsFileName or "results.txt"
    with open(results, "wb") as f:
        f.write(b"".join(ln + b"\n" for ln in res))
    with open(formats.nonmatch_name(results), "wb") as f:
        f.write(formats.nonmatch_fastq(m, seqs, counts, names))
    return st


def main(argv=None):
    ap = argparse.ArgumentParser(prog="muscato_b200.cli")
    ap.add_argument("config")
    ap.add_argument("--device", type=int, default=0)
    ap.add_argument("--from-fastq", action="store_true")
    a = ap.parse_args(argv)
    st = run(a.config, a.device, a.from_fastq)
    sys.stderr.write("muscato_b200: %d candidates, %d pairs, %d matches\n"
                     % (st["n_candidates"], st["n_pairs"], st["n_matches"]))


if __name__ == "__main__":
    main()
