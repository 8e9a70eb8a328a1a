"""Test-side helpers: building and running the CPU oracle (oracle/ is test infrastructure and
is only ever touched from tests/, smoke() and bench.py's baseline legs)."""
from __future__ import annotations

import json
import os
import subprocess
from typing import Dict, List, Optional

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_DIR = os.path.join(ROOT, "oracle")
ORACLE_BIN = os.path.join(ORACLE_DIR, "_build", "muscato_oracle")
GOLDEN = os.path.join(ROOT, "tests", "golden")


def build_oracle() -> str:
    src = os.path.join(ORACLE_DIR, "muscato_oracle.cc")
    if not os.path.exists(ORACLE_BIN) or os.path.getmtime(ORACLE_BIN) < os.path.getmtime(src):
        subprocess.run(["make", "-C", ORACLE_DIR], check=True, capture_output=True)
    return ORACLE_BIN


def run_oracle(args: List[str], **kw) -> subprocess.CompletedProcess:
    return subprocess.run([build_oracle()] + args, capture_output=True, text=True, **kw)


def oracle_prep_targets(src: str, seq_out: str, ids_out: str, rev: bool = False) -> None:
    a = ["prep_targets"] + (["-rev"] if rev else []) + [src, seq_out, ids_out]
    r = run_oracle(a)
    assert r.returncode == 0, r.stderr


def oracle_pipeline(workdir: str, fastq: str, genes_seq: str, genes_ids: str, cfg: Dict,
                    sub: str = "pipeline") -> Dict[str, str]:
    """Run the oracle on prepared targets; returns paths of its outputs."""
    tmp = os.path.join(workdir, "tmp")
    os.makedirs(tmp, exist_ok=True)
    c = dict(cfg)
    c.update(ReadFileName=fastq, GeneFileName=genes_seq, GeneIdFileName=genes_ids,
             ResultsFileName=os.path.join(workdir, "result.txt"), TempDir=tmp, SortPar=2, SortMem="5%")
    c.setdefault("Threads", 4)
    cpath = os.path.join(workdir, "oracle_config.json")
    with open(cpath, "w") as f:
        json.dump(c, f)
    r = run_oracle([sub, cpath])
    assert r.returncode == 0, (r.stdout, r.stderr)
    return {
        "config": cpath,
        "tmp": tmp,
        "results": c["ResultsFileName"],
        "nonmatch": os.path.join(workdir, "result.nonmatch.txt.fastq"),
        "matches": os.path.join(tmp, "matches.txt"),
        "reads_sorted": os.path.join(tmp, "reads_sorted.txt"),
        "stdout": r.stdout,
    }


def read_bytes(path: str) -> bytes:
    with open(path, "rb") as f:
        return f.read()


def read_lines(path: str) -> List[bytes]:
    d = read_bytes(path)
    lines = d.split(b"\n")
    if lines and lines[-1] == b"":
        lines.pop()
    return lines


def oracle_window_keys(tmp: str, reads: List[bytes], k: int) -> set:
    """win_<k>_sorted lines mapped back to read ids: {(k, read_id)}."""
    idx = {r: i for i, r in enumerate(reads)}
    out = set()
    for ln in read_lines(os.path.join(tmp, f"win_{k}_sorted.txt")):
        key, left, right = ln.split(b"\t")
        out.add((k, idx[left + key + right]))
    return out


def oracle_candidates(tmp: str, k: int) -> set:
    """{(k, gene, p)} for smatch_<k> lines whose k-mer is a key of win_<k>_sorted
    (the merge join of cmd/muscato_confirm/main.go:375-416 drops every other line)."""
    keys = {ln.split(b"\t")[0] for ln in read_lines(os.path.join(tmp, f"win_{k}_sorted.txt"))}
    out = set()
    for ln in read_lines(os.path.join(tmp, f"smatch_{k}.txt")):
        f = ln.split(b"\t")
        if f[0] in keys:
            out.add((k, int(f[3]), int(f[4])))
    return out


def random_dna(rng: np.random.Generator, n: int, alphabet: bytes = b"ACGT") -> bytes:
    a = np.frombuffer(alphabet, dtype=np.uint8)
    return a[rng.integers(0, len(a), size=n)].tobytes()


def write_case(workdir: str, raw_reads: List[bytes], names: Optional[List[bytes]], genes: List[bytes],
               gene_names: Optional[List[bytes]] = None):
    """Write reads.fastq + prepped target files (plain text) for the oracle."""
    os.makedirs(workdir, exist_ok=True)
    fq = os.path.join(workdir, "reads.fastq")
    with open(fq, "wb") as f:
        for i, r in enumerate(raw_reads):
            nm = names[i] if names else b"@read_%d" % i
            f.write(nm + b"\n" + r + b"\n+\n" + b"F" * len(r) + b"\n")
    gs = os.path.join(workdir, "genes_seq.txt")
    gi = os.path.join(workdir, "genes_ids.txt")
    with open(gs, "wb") as f:
        for g in genes:
            f.write(g + b"\n")
    with open(gi, "wb") as f:
        for i, g in enumerate(genes):
            nm = gene_names[i] if gene_names else b"gene_%d" % i
            f.write(b"%011d\t%s\t%d\n" % (i, nm, len(g)))
    return fq, gs, gi


# ---- second, independent restatement of the path: the closed form of SURVEY.md App. A --------------
# Plain string search per read; used to cross-check the oracle on the CPU (tests/test_oracle_closed_form.py)
# and the CUDA path at full size (tests/test_gpu_fullsize.py).  `cfg` is a muscato_b200.config.Config.
def count_dinuc(s: bytes) -> int:
    return len({s[i:i + 2] for i in range(len(s) - 1)})


def closed_form_matches(read: bytes, tgt: bytes, offs: np.ndarray, cfg: Config):
    """All (gene, pos, nx) the reference produces for one read before the MMTol rule (App. A.2)."""
    W, MRL, L = cfg.WindowWidth, cfg.MaxReadLength, len(read)
    nmiss = cfg.nmiss(L)
    out = {}
    r = np.frombuffer(read, dtype=np.uint8)
    for q1 in cfg.Windows:
        q2 = q1 + W
        if L < q2 or count_dinuc(read[q1:q2]) < cfg.MinDinuc:
            continue
        key = read[q1:q2]
        at = tgt.find(key)
        while at >= 0:
            g = int(np.searchsorted(offs, at, side="right") - 1)
            goff, glen = int(offs[g]), int(offs[g + 1] - offs[g])
            p = at - goff
            pos = p - q1
            ok = p + W <= glen and pos >= 0
            if ok:
                if p == 0:
                    ok = L <= min(100 - W, glen)                      # the literal 100 (Q1)
                else:
                    ok = L - q2 <= min(p + W + MRL - q2, glen) - (p + W)
            if ok:
                t = np.frombuffer(tgt[goff + pos:goff + pos + L], dtype=np.uint8)
                nx = int((t != r).sum())
                if nx <= nmiss:
                    out[(g, pos)] = nx
            at = tgt.find(key, at + 1)
    return out


def sequential_restatement(seqs, genes, cfg):
    """Third restatement, this time of the reference's SEQUENTIAL structure (needed for MaxMatches, Q7,
    which the closed form cannot express): per window the read records `key left right`
    (cmd/muscato_window_reads/main.go:100-126) and the candidate records `key left right %011d pos`
    (cmd/muscato_screen/main.go:294-365, incl. the position-0 record with its literal 100) in bytewise
    line order (`LC_ALL=C sort`, cmd/muscato/main.go:237-385), the candidate-major / read-minor double
    loop of searchpairs with its bounded result list (cmd/muscato_confirm/main.go:171-250, qinsert
    :424-448), the union with exact de-duplication and the MMTol rule
    (cmd/muscato_combine_windows/main.go:36-60).  Returns the set of matches.txt lines.  Pure Python,
    small cases only."""
    W, MRL, MM = cfg.WindowWidth, cfg.MaxReadLength, int(cfg.MaxMatches)
    first = cfg.MatchMode == "first"
    lines = set()
    for q1 in cfg.Windows:
        q2 = q1 + W
        readrecs = {}
        for r in seqs:
            if len(r) >= q2 and (cfg.MinDinuc <= 0 or count_dinuc(r[q1:q2]) >= cfg.MinDinuc):
                readrecs.setdefault(r[q1:q2], []).append((r[:q1], r[q2:]))
        candrecs = {}
        for g, t in enumerate(genes):
            for p in range(0, len(t) - W + 1):
                key = t[p:p + W]
                if key not in readrecs:              # the merge join drops every other candidate
                    continue
                if p == 0:
                    if q1 != 0:
                        continue
                    left, right = b"", t[W:min(100 - q2, len(t))]
                else:
                    if p - q1 < 0:
                        continue
                    left, right = t[p - q1:p], t[p + W:min(p + W + MRL - q2, len(t))]
                candrecs.setdefault(key, []).append((left, right, b"%011d" % g, b"%d" % p))
        for key, rr in readrecs.items():
            rr = sorted(rr, key=lambda x: x[0] + b"\t" + x[1])
            cc = sorted(candrecs.get(key, []), key=lambda x: b"\t".join(x))
            kept = []                                # (nx, line)
            stop = False
            for ml, mr, mg, mp in cc:
                for sl, sr in rr:
                    if len(sr) > len(mr):
                        continue
                    nx = sum(a != b for a, b in zip(ml, sl)) + sum(a != b for a, b in zip(mr[:len(sr)], sr))
                    if nx > int((1 - cfg.PMatch) * float(W + len(sl) + len(sr))):
                        continue
                    line = b"%s\t%s\t%d\t%d\t%s" % (sl + key + sr, ml + key + mr[:len(sr)], int(mp) - len(ml), nx, mg)
                    if first:
                        kept.append((nx, line))
                        if len(kept) > MM:
                            stop = True
                            break
                    else:
                        kept.append((nx, line))
                        i = len(kept) - 1
                        while i > 0:
                            j = (i - 1) // 2
                            if kept[j][0] > kept[i][0]:
                                kept[j], kept[i] = kept[i], kept[j]
                                i = j
                            else:
                                break
                        if len(kept) > MM:
                            del kept[MM:]
                if stop:
                    break
            lines.update(ln for _, ln in kept)
    best = {}
    for ln in lines:
        f = ln.split(b"\t")
        best[f[0]] = min(best.get(f[0], 1 << 30), int(f[3]))
    return {ln for ln in lines if int(ln.split(b"\t")[3]) <= best[ln.split(b"\t")[0]] + cfg.MMTol}
