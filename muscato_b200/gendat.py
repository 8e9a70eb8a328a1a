"""Synthetic inputs with the shape of muscato_gendat (cmd/muscato_gendat/main.go:39-136):
iid uniform A/T/G/C reads and genes; for gene i < NumGene/2, read (i % 10) is copied into
the gene at offset (i % 10) (:122-125).  Go's unseeded math/rand stream cannot be
reproduced, so a seeded numpy generator is used and the seed is part of the workload name.

`mutated_fraction` adds the S1m variant of BASELINE.md: that fraction of the reads is
sampled from uniform random target positions with per-base substitution probability
`sub_rate` (Binomial(L, sub_rate) substitutions), so that confirm sees 0..5 mismatches."""
from __future__ import annotations

import dataclasses
from typing import List, Tuple

import numpy as np

_BASES = np.frombuffer(b"ATGC", dtype=np.uint8)  # gendat order (:83)
_COMP = np.zeros(256, dtype=np.uint8)
for _a, _b in zip(b"ATGCX", b"TACGX"):
    _COMP[_a] = _b


@dataclasses.dataclass
class Synthetic:
    read_ascii: np.ndarray   # uint8, all unique reads concatenated in reads_sorted (bytewise) order
    read_offs: np.ndarray    # uint64 [U+1]
    read_counts: np.ndarray  # multiplicity of each unique read
    read_first: np.ndarray   # index of the first raw read carrying each unique sequence
    target_ascii: np.ndarray
    target_offs: np.ndarray  # uint64 [G+1]
    n_raw_reads: int
    read_len: int

    @property
    def n_reads(self) -> int:
        return len(self.read_offs) - 1

    @property
    def n_targets(self) -> int:
        return len(self.target_offs) - 1

    @property
    def target_bases(self) -> int:
        return int(self.target_offs[-1])

    def reads_list(self) -> List[bytes]:
        b = self.read_ascii.tobytes()
        o = self.read_offs
        return [b[int(o[i]):int(o[i + 1])] for i in range(self.n_reads)]

    def targets_list(self) -> List[bytes]:
        b = self.target_ascii.tobytes()
        o = self.target_offs
        return [b[int(o[i]):int(o[i + 1])] for i in range(self.n_targets)]


def revcomp_rows(genes: np.ndarray) -> np.ndarray:
    """revcomp (cmd/muscato_prep_targets/main.go:48-66) on a [G, len] uint8 matrix."""
    return _COMP[genes[:, ::-1]]


def generate(num_read: int, read_len: int, num_gene: int, gene_len: int, seed: int = 1, rev: bool = False,
             mutated_fraction: float = 0.0, sub_rate: float = 0.02, n_shards: int = 1) -> Synthetic:
    """num_gene genes in total.  With n_shards > 1 the gene list is n_shards consecutive gendat
    databases of num_gene/n_shards genes each (every one plants the ten reads in its own first
    half), so that contiguous target sharding gives every rank a statistically identical shard."""
    if num_read < 10:
        raise ValueError("numRead must be at least 10")  # :148-150
    rng = np.random.default_rng(seed)
    reads = _BASES[rng.integers(0, 4, size=(num_read, read_len), dtype=np.uint8)]
    genes = _BASES[rng.integers(0, 4, size=(num_gene, gene_len), dtype=np.uint8)]
    # plant read i%10 at offset i%10 of gene i for i < NumGene/2 (:122-125)
    per = max(1, num_gene // max(1, n_shards))
    for i in range(num_gene):
        li = i % per
        if li >= per // 2:
            continue
        j = li % 10
        n = min(read_len, gene_len - j)
        if n > 0:
            genes[i, j:j + n] = reads[j, :n]
    if mutated_fraction > 0 and gene_len >= read_len:
        nm = int(num_read * mutated_fraction)
        idx = rng.choice(np.arange(10, num_read), size=min(nm, num_read - 10), replace=False)
        g = rng.integers(0, num_gene, size=len(idx))
        p = rng.integers(0, gene_len - read_len + 1, size=len(idx))
        cols = p[:, None] + np.arange(read_len)[None, :]
        sampled = genes[g[:, None], cols]
        mut = rng.random(size=sampled.shape) < sub_rate
        shift = rng.integers(1, 4, size=sampled.shape)
        # substitute with a *different* base
        idx_of = np.zeros(256, dtype=np.int64)
        idx_of[_BASES] = np.arange(4)
        new_idx = (idx_of[sampled] + shift) % 4
        sampled = np.where(mut, _BASES[new_idx], sampled)
        reads[idx] = sampled
    if rev:
        rc = revcomp_rows(genes)
        both = np.empty((2 * num_gene, gene_len), dtype=np.uint8)
        both[0::2] = genes   # 2i = forward, 2i+1 = reverse complement (:115-134)
        both[1::2] = rc
        genes = both
    # uniqify in bytewise order (prepReads: sort | muscato_uniqify)
    view = np.ascontiguousarray(reads).view(np.dtype((np.void, read_len))).ravel()
    uniq, first, counts = np.unique(view, return_index=True, return_counts=True)
    ureads = uniq.view(np.uint8).reshape(-1, read_len)
    U = ureads.shape[0]
    G = genes.shape[0]
    return Synthetic(
        read_ascii=np.ascontiguousarray(ureads).ravel(),
        read_offs=(np.arange(U + 1, dtype=np.uint64) * np.uint64(read_len)),
        read_counts=counts.astype(np.int64),
        read_first=first.astype(np.int64),
        target_ascii=np.ascontiguousarray(genes).ravel(),
        target_offs=(np.arange(G + 1, dtype=np.uint64) * np.uint64(gene_len)),
        n_raw_reads=num_read,
        read_len=read_len,
    )


def write_oracle_inputs(syn: Synthetic, fastq_path: str, genes_path: str, ids_path: str) -> None:
    """Materialise a Synthetic as the files the oracle pipeline reads: a fastq with one record per
    *unique* read carrying its multiplicity as repeated records, the prepped target file (one
    sequence per line) and the gene id file (`%011d\\tname\\tlen`)."""
    reads = syn.reads_list()
    with open(fastq_path, "wb") as f:
        for i, r in enumerate(reads):
            for c in range(int(syn.read_counts[i])):
                f.write(b"read_%d_%d\n%s\n+\n%s\n" % (i, c, r, b"!" * len(r)))
    tg = syn.targets_list()
    with open(genes_path, "wb") as f:
        for t in tg:
            f.write(t + b"\n")
    with open(ids_path, "wb") as f:
        for i, t in enumerate(tg):
            f.write(b"%011d\tgene_%d\t%d\n" % (i, i, len(t)))


# ---------------------------------------------------------------------------------------------
# Block-structured generator for the large configurations (csrc/host/gendat.cc): S2 / S3 / S4.
# ---------------------------------------------------------------------------------------------
@dataclasses.dataclass(frozen=True)
class BlockSpec:
    """A workload of `n_blocks` independent blocks; block b depends on (seed, b) only.  Reads of a
    block are `planted_per_block` copies (with sub256/256 substitutions per base) of windows of
    the block's own targets followed by iid uniform reads."""
    seed: int
    n_blocks: int
    reads_per_block: int
    read_len: int
    planted_per_block: int
    sub256: int
    genes_per_block: int
    gene_len: int
    rev: bool = True
    period: int = 0           # > 0: tandem-repeat targets with units of period_min..period bases (S4)
    period_min: int = 1
    target_sub256: int = 0    # substitutions inside tandem-repeat targets

    @property
    def targets_per_block(self) -> int:
        return self.genes_per_block * (2 if self.rev else 1)

    @property
    def n_reads(self) -> int:
        return self.n_blocks * self.reads_per_block

    @property
    def n_targets(self) -> int:
        return self.n_blocks * self.targets_per_block

    @property
    def target_bases(self) -> int:
        return self.n_targets * self.gene_len

    def prefix(self, n_blocks: int) -> "BlockSpec":
        return dataclasses.replace(self, n_blocks=n_blocks)


_gen_lib = None


def _gendat_lib():
    global _gen_lib
    if _gen_lib is None:
        import ctypes as C
        import os
        from . import build as _b
        if not os.path.exists(_b.GENDAT_PATH):
            _b.build_gendat()
        lib = C.CDLL(_b.GENDAT_PATH)
        lib.msc_gen_targets.argtypes = [C.c_uint64, C.c_uint64, C.c_uint32, C.c_uint32, C.c_int, C.c_uint32, C.c_uint32,
                                        C.c_void_p]
        lib.msc_gen_targets.restype = None
        lib.msc_gen_reads.argtypes = [C.c_uint64, C.c_uint64, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32, C.c_void_p,
                                      C.c_uint32, C.c_uint32, C.c_void_p, C.c_void_p, C.c_void_p]
        lib.msc_gen_reads.restype = None
        _gen_lib = lib
    return _gen_lib


def generate_blocks(spec: BlockSpec, blocks=None, reads_out: np.ndarray = None, targets_out: np.ndarray = None,
                    want_reads: bool = True, want_targets: bool = True, want_plants: bool = False, threads: int = None):
    """Generate blocks `blocks` (default: all) of `spec` on all host threads.  Returns
    (reads uint8 [n*rpb*L] or None, targets uint8 [n*tpb*GL] or None, plants or None) where plants =
    (gene int64 [n*planted] as GLOBAL-in-this-selection target index, pos int32).  reads_out /
    targets_out may be preallocated (e.g. pinned) uint8 buffers of the right size."""
    import os
    from concurrent.futures import ThreadPoolExecutor
    lib = _gendat_lib()
    blocks = list(range(spec.n_blocks)) if blocks is None else list(blocks)
    nb = len(blocks)
    rbytes = spec.reads_per_block * spec.read_len
    tbytes = spec.targets_per_block * spec.gene_len
    reads = targets = None
    if want_reads:
        reads = reads_out if reads_out is not None else np.empty(nb * rbytes, dtype=np.uint8)
        assert reads.dtype == np.uint8 and reads.size == nb * rbytes and reads.flags.c_contiguous
    if want_targets:
        targets = targets_out if targets_out is not None else np.empty(nb * tbytes, dtype=np.uint8)
        assert targets.dtype == np.uint8 and targets.size == nb * tbytes and targets.flags.c_contiguous
    pg = pp = None
    if want_plants and want_reads:
        pg = np.zeros(nb * spec.planted_per_block, dtype=np.int32)
        pp = np.zeros(nb * spec.planted_per_block, dtype=np.int32)

    def one(i):
        b = blocks[i]
        if want_targets:
            tv = targets[i * tbytes:(i + 1) * tbytes]
        else:
            tv = np.empty(tbytes, dtype=np.uint8)
        if want_targets or want_reads:
            lib.msc_gen_targets(spec.seed, b, spec.genes_per_block, spec.gene_len, int(spec.rev), spec.period | (spec.period_min << 8) if spec.period else 0,
                                spec.target_sub256, tv.ctypes.data)
        if want_reads:
            rv = reads[i * rbytes:(i + 1) * rbytes]
            n_pl = spec.planted_per_block
            lib.msc_gen_reads(spec.seed, b, spec.reads_per_block, spec.read_len, n_pl, spec.sub256, tv.ctypes.data,
                              spec.targets_per_block, spec.gene_len, rv.ctypes.data,
                              pg[i * n_pl:].ctypes.data if pg is not None and n_pl else None,
                              pp[i * n_pl:].ctypes.data if pp is not None and n_pl else None)

    threads = threads or min(32, os.cpu_count() or 1)
    if nb <= 1 or threads <= 1:
        for i in range(nb):
            one(i)
    else:
        with ThreadPoolExecutor(max_workers=threads) as ex:
            list(ex.map(one, range(nb)))
    plants = None
    if pg is not None:
        g = pg.astype(np.int64) + np.repeat(np.arange(nb, dtype=np.int64) * spec.targets_per_block, spec.planted_per_block)
        plants = (g, pp)
    return reads, targets, plants
