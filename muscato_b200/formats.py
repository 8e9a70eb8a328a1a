"""File formats either side of the hot path and the text epilogue that
turns integer matches into matches.txt / results.txt / the non-match fastq.

Every function cites the reference code whose output format it reproduces.  Orderings are
bytewise (the driver forces LC_ALL=C, cmd/muscato/main.go:906-912)."""
from __future__ import annotations

from typing import List, Sequence, Tuple

import numpy as np

from . import sz

_WS = b" \t\n\v\f\r"


def _fields(line: bytes) -> List[bytes]:
    return line.split()  # bytes.Fields on ASCII input


def parse_reads_sorted(data: bytes) -> Tuple[List[bytes], List[bytes], List[bytes]]:
    """reads_sorted.txt.sz lines `seq\\tcount\\tnames` (cmd/muscato_uniqify/main.go:89-110).
    Returns (seqs, counts, names); names is the raw third tab field."""
    seqs, counts, names = [], [], []
    for line in data.split(b"\n"):
        if not line:
            continue
        t = line.split(b"\t", 2)
        seqs.append(t[0])
        counts.append(t[1] if len(t) > 1 else b"")
        names.append(t[2] if len(t) > 2 else b"")
    return seqs, counts, names


def parse_targets(data: bytes) -> List[bytes]:
    """Target file: one sequence per line; text before the first tab (cmd/muscato_screen/main.go:446-449)."""
    lines = data.split(b"\n")
    if lines and lines[-1] == b"":
        lines.pop()
    return [ln.rstrip(b"\r").split(b"\t", 1)[0] for ln in lines]


def parse_gene_ids(data: bytes) -> Tuple[List[bytes], List[bytes]]:
    """Gene id file lines `%011d\\tname\\tlen` (cmd/muscato_prep_targets/main.go:123-132, 165)."""
    names, lens = [], []
    for line in data.split(b"\n"):
        if not line:
            continue
        t = line.split(b"\t")
        names.append(t[1])
        lens.append(t[2] if len(t) > 2 else b"")
    return names, lens


def load_reads_sorted(path: str):
    return parse_reads_sorted(sz.read_file(path))


def load_targets(path: str):
    return parse_targets(sz.read_file(path))


def load_gene_ids(path: str):
    return parse_gene_ids(sz.read_file(path))


def matches_lines(matches: np.ndarray, reads: Sequence[bytes], targets: Sequence[bytes]) -> List[bytes]:
    """matches.txt.sz content: `read\\ttarget[pos:pos+L]\\tpos\\tnx\\t%011d(gene)`
    (cmd/muscato_confirm/main.go:221-230), whole-line sorted as `sort -u` leaves them
    (cmd/muscato/main.go:453-463)."""
    out = []
    for m in matches:
        r = reads[int(m["read_id"])]
        pos = int(m["pos"])
        t = targets[int(m["gene_id"])][pos:pos + len(r)]
        out.append(b"%s\t%s\t%d\t%d\t%011d" % (r, t, pos, int(m["nx"]), int(m["gene_id"])))
    out.sort()
    return out


def results_lines(matches: np.ndarray, reads: Sequence[bytes], counts: Sequence[bytes], names: Sequence[bytes],
                  targets: Sequence[bytes], gene_names: Sequence[bytes], gene_lens: Sequence[bytes]) -> List[bytes]:
    """results.txt: sortByGeneId + joinGeneNames + `sort -k1` + joinReadNames
    (cmd/muscato/main.go:507-676).  Columns: read, target subsequence, pos, nx, gene name,
    gene length, read count, read names; ordered bytewise on the first six columns."""
    rows = []
    for m in matches:
        rid = int(m["read_id"])
        g = int(m["gene_id"])
        r = reads[rid]
        pos = int(m["pos"])
        t = targets[g][pos:pos + len(r)]
        six = b"%s\t%s\t%d\t%d\t%s\t%s" % (r, t, pos, int(m["nx"]), gene_names[g], gene_lens[g])
        rows.append((six, rid))
    rows.sort(key=lambda x: x[0])
    return [six + b"\t" + counts[rid] + b"\t" + names[rid] for six, rid in rows]


def nonmatch_name(results_path: str) -> str:
    """Output name rule of cmd/muscato_nonmatch/main.go:66-71."""
    if "/" in results_path:
        a, b = results_path.rsplit("/", 1)
        a += "/"
    else:
        a, b = "", results_path
    c = b.split(".")
    d = c[-1]
    c[-1] = "nonmatch"
    c.append(d + ".fastq")
    return a + ".".join(c)


def nonmatch_fastq(matches: np.ndarray, reads: Sequence[bytes], counts: Sequence[bytes],
                   names: Sequence[bytes]) -> bytes:
    """cmd/muscato_nonmatch/main.go:95-113 with an exact matched-read set (Q10)."""
    matched = np.zeros(len(reads), dtype=bool)
    if len(matches):
        matched[matches["read_id"]] = True
    out = []
    for i, r in enumerate(reads):
        if matched[i]:
            continue
        nm = _fields(names[i])
        first = nm[0] if nm else b""
        out.append(first + b"#" + counts[i] + b"\n" + r + b"\n+\n" + b"!" * len(r) + b"\n")
    return b"".join(out)


def nonmatch_fastq_from_ids(ids, reads: Sequence[bytes], counts: Sequence[bytes], names: Sequence[bytes]) -> bytes:
    """Same file as nonmatch_fastq(), from the unmatched read ids the device reports
    (msc_fetch_nonmatch) instead of a scan over the matches."""
    out = []
    for i in ids:
        i = int(i)
        nm = _fields(names[i])
        first = nm[0] if nm else b""
        out.append(first + b"#" + counts[i] + b"\n" + reads[i] + b"\n+\n" + b"!" * len(reads[i]) + b"\n")
    return b"".join(out)


def prep_reads_uniqify(fastq: bytes, min_len: int, max_len: int):
    """prepReads = muscato_prep_reads | sort | muscato_uniqify (cmd/muscato/main.go:152-221;
    cmd/muscato_prep_reads/main.go:46-92; cmd/muscato_uniqify/main.go:77-135).
    Returns (seqs, counts, names) in reads_sorted order."""
    lines = fastq.split(b"\n")
    if lines and lines[-1] == b"":
        lines.pop()
    recs = []
    for i in range(0, len(lines) - len(lines) % 4, 4):
        name = lines[i].rstrip(b"\r")
        seq = lines[i + 1].rstrip(b"\r")
        if len(seq) < min_len:
            continue
        seq = bytes(c if c in b"ATCG" else 0x58 for c in seq)
        if len(seq) > max_len:
            seq = seq[:max_len]
        if len(name) > 1000:
            name = name[:995] + b"..."
        recs.append(seq + b"\t" + name)
    recs.sort()
    seqs, counts, names = [], [], []
    cur, cur_names = None, []

    def flush():
        na = b";".join(cur_names)
        if len(na) > 1000:
            na = na[:996] + b"..."
        seqs.append(cur)
        counts.append(b"%d" % len(cur_names))
        names.append(na)

    for rec in recs:
        t = rec.split(b"\t")
        if t[0] != cur:
            if cur is not None:
                flush()
            cur, cur_names = t[0], []
        cur_names.append(t[1])
    if cur is not None:
        flush()
    return seqs, counts, names


def parse_fastq(fastq: bytes):
    """(names, sequences) of a fastq file as the reference reads it: 4-line records, the header and
    sequence lines right-trimmed of '\\r' (utils/fastq.go:40-70; cmd/muscato_prep_reads/main.go:46-58)."""
    lines = fastq.split(b"\n")
    if lines and lines[-1] == b"":
        lines.pop()
    names, seqs = [], []
    for i in range(0, len(lines) - len(lines) % 4, 4):
        names.append(lines[i].rstrip(b"\r"))
        seqs.append(lines[i + 1].rstrip(b"\r"))
    return names, seqs


def uniqify_from_groups(raw_names: Sequence[bytes], raw_seqs: Sequence[bytes], perm, group_start, max_len: int):
    """Host half of the device-side prepReads (msc_prep_reads): the device returns the sorted
    permutation of the raw reads and the group boundaries; counts and names are joined here.
    Inside a group the `seq\\tname` lines are in bytewise name order (cmd/muscato/main.go:183-199),
    the name column is the text before its first tab (cmd/muscato_uniqify/main.go:100-110) and the
    joined names are cut to 996 bytes + "..." (:89-93).  Returns (seqs, counts, names)."""
    seqs, counts, names = [], [], []
    for u in range(len(group_start) - 1):
        members = [int(perm[j]) for j in range(int(group_start[u]), int(group_start[u + 1]))]
        nm = []
        for r in members:
            n = raw_names[r]
            if len(n) > 1000:
                n = n[:995] + b"..."
            nm.append(n)
        nm.sort()
        na = b";".join(n.split(b"\t")[0] for n in nm)
        if len(na) > 1000:
            na = na[:996] + b"..."
        s = raw_seqs[members[0]][:max_len]
        seqs.append(bytes(c if c in b"ATCG" else 0x58 for c in s))
        counts.append(b"%d" % len(members))
        names.append(na)
    return seqs, counts, names
