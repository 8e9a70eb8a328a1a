"""Host-side handle on the CUDA hot path: a thin object over the C ABI
(include/muscato_b200.h).  All computation happens in libmuscato_b200.so on the GPU."""
from __future__ import annotations

import ctypes as C
import threading
from typing import Iterable, Sequence, Tuple, Union

import numpy as np

from . import _capi
from .config import Config

MATCH_DTYPE = np.dtype([("read_id", "<u4"), ("gene_id", "<u4"), ("pos", "<u4"), ("nx", "<u4")])
KEY_DTYPE = np.dtype([("window", "<u4"), ("read_id", "<u4")])
CAND_DTYPE = np.dtype([("gene_id", "<u4"), ("p", "<u4"), ("read_id", "<u4"), ("window", "<u4")])

SeqInput = Union[Sequence[bytes], Tuple[np.ndarray, np.ndarray]]


class MuscatoError(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__(f"muscato_b200 error {code}: {msg}")
        self.code = code


def concat_sequences(seqs: Iterable[bytes]) -> Tuple[np.ndarray, np.ndarray]:
    """list of byte strings -> (ascii uint8 array, uint64 offsets[n+1])."""
    seqs = list(seqs)
    offs = np.zeros(len(seqs) + 1, dtype=np.uint64)
    if seqs:
        offs[1:] = np.cumsum([len(s) for s in seqs], dtype=np.uint64)
    ascii_ = np.frombuffer(b"".join(seqs), dtype=np.uint8)
    return ascii_, offs


class _DevArray:
    """Minimal __cuda_array_interface__ holder so torch can view a library-owned device buffer."""

    def __init__(self, ptr: int, n: int, typestr: str):
        self.__cuda_array_interface__ = {"shape": (n,), "typestr": typestr, "data": (ptr, False), "version": 2}


class HotPath:
    def __init__(self, cfg: Config, device: int = 0, keep_ascii: bool = False, bloom_bits_per_key: int = 0):
        self._lib = _capi.load()
        self.cfg = cfg
        mc = cfg.to_msc(device=device, keep_ascii=keep_ascii, bloom_bits_per_key=bloom_bits_per_key)
        err = C.create_string_buffer(512)
        self._ctx = self._lib.msc_create(C.byref(mc), err, 512)
        if not self._ctx:
            raise MuscatoError(-1, err.value.decode(errors="replace"))
        self.n_reads = 0
        self.n_targets = 0
        self.n_shards = 1
        self._keep = []

    # -- lifecycle -------------------------------------------------------------------
    def close(self):
        if getattr(self, "_ctx", None):
            self._lib.msc_destroy(self._ctx)
            self._ctx = None

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, rc: int):
        if rc != 0:
            raise MuscatoError(rc, self._lib.msc_last_error(self._ctx).decode(errors="replace"))

    @staticmethod
    def _as_arrays(seqs: SeqInput) -> Tuple[np.ndarray, np.ndarray]:
        if isinstance(seqs, tuple) and len(seqs) == 2 and isinstance(seqs[0], np.ndarray):
            a, o = seqs
            return np.ascontiguousarray(a, dtype=np.uint8), np.ascontiguousarray(o, dtype=np.uint64)
        return concat_sequences(seqs)

    # -- the path -------------------------------------------------------------------
    def set_reads(self, seqs: SeqInput):
        a, o = self._as_arrays(seqs)
        self.n_reads = len(o) - 1
        self._check(self._lib.msc_set_reads(self._ctx, a.ctypes.data, o.ctypes.data, self.n_reads))

    def prep_reads(self, raw_seqs: SeqInput, min_read_length: int = 0):
        """Device-side prepReads (prep_reads | sort | uniqify of the sequence column,
        cmd/muscato/main.go:152-221): sorts and collapses the raw fastq sequences on the GPU and
        installs the unique reads as the read set.  Returns (n_kept, n_unique)."""
        a, o = self._as_arrays(raw_seqs)
        kept, uniq = C.c_uint64(0), C.c_uint64(0)
        self._check(self._lib.msc_prep_reads(self._ctx, a.ctypes.data, o.ctypes.data, len(o) - 1, int(min_read_length),
                                             C.byref(kept), C.byref(uniq)))
        self.n_reads = int(uniq.value)
        self._n_kept = int(kept.value)
        return self._n_kept, self.n_reads

    def read_groups(self):
        """(perm, group_start) of the last prep_reads: perm = raw read indices in sorted order,
        unique read u = perm[group_start[u]:group_start[u + 1]]."""
        perm = np.empty(max(1, self._n_kept), dtype=np.uint32)
        gs = np.empty(self.n_reads + 1, dtype=np.uint32)
        self._check(self._lib.msc_fetch_read_groups(self._ctx, perm.ctypes.data, gs.ctypes.data))
        return perm[: self._n_kept], gs

    def unique_reads(self):
        """The unique reads of the last prep_reads as a list of bytes (needs keep_ascii)."""
        nb = int(self._lib.msc_unique_reads_bytes(self._ctx))
        a = np.empty(max(1, nb), dtype=np.uint8)
        o = np.empty(self.n_reads + 1, dtype=np.uint64)
        self._check(self._lib.msc_fetch_unique_reads(self._ctx, a.ctypes.data, o.ctypes.data))
        buf = a.tobytes()
        return [buf[int(o[i]):int(o[i + 1])] for i in range(self.n_reads)]

    def set_targets(self, seqs: SeqInput):
        a, o = self._as_arrays(seqs)
        self.n_targets = len(o) - 1
        self._check(self._lib.msc_set_targets(self._ctx, a.ctypes.data, o.ctypes.data, self.n_targets))

    def set_targets_packed(self, words: np.ndarray, xplane, offs: np.ndarray):
        """Targets that are already 2-bit packed (the persistent target cache): words / xplane are uint64
        arrays of (total bases + 31) // 32 entries in the layout of include/muscato_b200.h; xplane may
        be None when no target contains X."""
        words = np.ascontiguousarray(words, dtype=np.uint64)
        offs = np.ascontiguousarray(offs, dtype=np.uint64)
        xp = None if xplane is None else np.ascontiguousarray(xplane, dtype=np.uint64)
        self.n_targets = len(offs) - 1
        self._check(self._lib.msc_set_targets_packed(self._ctx, words.ctypes.data, xp.ctypes.data if xp is not None else None,
                                                     offs.ctypes.data, self.n_targets))

    def fetch_packed_targets(self):
        """(words, xplane, has_x): the packed form of the current targets."""
        n = int(self._lib.msc_packed_target_words(self._ctx))
        words = np.zeros(max(1, n), dtype=np.uint64)
        xp = np.zeros(max(1, n), dtype=np.uint64)
        hx = C.c_int32(0)
        self._check(self._lib.msc_fetch_packed_targets(self._ctx, words.ctypes.data, xp.ctypes.data, C.byref(hx)))
        return words[:n], xp[:n], bool(hx.value)

    def set_reads_ptr(self, ascii_ptr: int, offs_ptr: int, n: int):
        """Raw-pointer variant (e.g. pinned host buffers owned by the caller)."""
        self.n_reads = n
        self._check(self._lib.msc_set_reads(self._ctx, ascii_ptr, offs_ptr, n))

    def set_reads_device(self, d_ascii_ptr: int, d_offs_ptr: int, n: int, total_bytes: int):
        """Reads already resident on this context's device (e.g. after an NCCL broadcast)."""
        self.n_reads = n
        self._check(self._lib.msc_set_reads_device(self._ctx, d_ascii_ptr, d_offs_ptr, n, total_bytes))

    def set_targets_ptr(self, ascii_ptr: int, offs_ptr: int, n: int):
        self.n_targets = n
        self._check(self._lib.msc_set_targets(self._ctx, ascii_ptr, offs_ptr, n))

    def set_targets_from(self, other: "HotPath"):
        """The packed targets of another context on the same device, copied device to device (msc_set_targets_from)."""
        self.n_targets = other.n_targets
        self._check(self._lib.msc_set_targets_from(self._ctx, other._ctx))

    def rebuild(self, what: int = 3):
        self._check(self._lib.msc_rebuild(self._ctx, what))

    def screen(self):
        self._check(self._lib.msc_screen(self._ctx))

    def confirm(self):
        self._check(self._lib.msc_confirm(self._ctx))

    def combine(self):
        self._check(self._lib.msc_combine(self._ctx))

    def run(self):
        self._check(self._lib.msc_run(self._ctx))

    def run_stages(self, rebuild_what: int, stages: int):
        self._check(self._lib.msc_run_stages(self._ctx, rebuild_what, stages))

    def rebuild_and_run(self, what: int = 3):
        self._check(self._lib.msc_rebuild_and_run(self._ctx, what))

    def stream(self) -> int:
        """The cudaStream_t (as an integer) every kernel of this context runs on."""
        return int(self._lib.msc_stream(self._ctx) or 0)

    def best_device(self) -> _DevArray:
        ptr = self._lib.msc_best_device(self._ctx)
        if not ptr:
            raise MuscatoError(_capi.MSC_ERR_STATE, "best array not available (run confirm first)")
        # with sharded targets element [n_reads] carries the MaxMatches flag through the all-reduce
        return _DevArray(ptr, self.n_reads + (1 if self.n_shards > 1 else 0), "<i4")

    # -- MaxMatches across target shards (include/muscato_b200.h, msc_set_shards) --------------
    def set_shards(self, n_shards: int):
        self._check(self._lib.msc_set_shards(self._ctx, int(n_shards)))
        self.n_shards = int(n_shards)

    def shard_overflow(self) -> bool:
        """After combine: did any shard flag a key group with more than MaxMatches / n_shards passing pairs?"""
        return bool(self._lib.msc_shard_overflow(self._ctx))

    def overflow_keys(self) -> np.ndarray:
        out = C.POINTER(C.c_uint64)()
        n = C.c_uint64(0)
        self._check(self._lib.msc_overflow_keys(self._ctx, C.byref(out), C.byref(n)))
        try:
            return np.ctypeslib.as_array(out, shape=(n.value,)).copy() if n.value else np.zeros(0, dtype=np.uint64)
        finally:
            self._lib.msc_free(out)

    def diverted_record_bytes(self) -> int:
        return int(self._lib.msc_diverted_record_bytes(self._ctx))

    def divert_groups(self, keys: np.ndarray, gene_base: int) -> np.ndarray:
        """uint8 [n, record_bytes]: the passing pairs of the flagged key groups of this shard."""
        keys = np.ascontiguousarray(keys, dtype=np.uint64)
        out = C.POINTER(C.c_uint8)()
        n = C.c_uint64(0)
        self._check(self._lib.msc_divert_groups(self._ctx, keys.ctypes.data, len(keys), int(gene_base), C.byref(out),
                                                C.byref(n)))
        rb = self.diverted_record_bytes()
        try:
            if n.value == 0:
                return np.zeros((0, rb), dtype=np.uint8)
            return np.ctypeslib.as_array(out, shape=(n.value * rb,)).copy().reshape(-1, rb)
        finally:
            self._lib.msc_free(out)

    def replay_diverted(self, recs: np.ndarray) -> np.ndarray:
        """The reference's sequential MaxMatches truncation over the diverted pairs of ALL shards;
        returns the survivors (global gene ids), de-duplicated, ordered by (read, gene, pos)."""
        recs = np.ascontiguousarray(recs, dtype=np.uint8)
        rb = self.diverted_record_bytes()
        assert recs.size % rb == 0
        out = C.POINTER(_capi.msc_match)()
        n = C.c_uint64(0)
        self._check(self._lib.msc_replay_diverted(self._ctx, recs.ctypes.data, recs.size // rb, C.byref(out), C.byref(n)))
        try:
            if n.value == 0:
                return np.zeros(0, dtype=MATCH_DTYPE)
            return np.frombuffer(C.string_at(out, n.value * C.sizeof(_capi.msc_match)), dtype=MATCH_DTYPE).copy()
        finally:
            self._lib.msc_free(out)

    def matches_device(self):
        """(holder, n): device view of the combined matches as int32[n*4] (read, gene, pos, nx)."""
        n = C.c_uint64(0)
        ptr = self._lib.msc_matches_device(self._ctx, C.byref(n))
        if not ptr:
            raise MuscatoError(_capi.MSC_ERR_STATE, "matches not available (run combine first)")
        return _DevArray(ptr, int(n.value) * 4, "<i4"), int(n.value)

    def fetch(self) -> np.ndarray:
        out = C.POINTER(_capi.msc_match)()
        n = C.c_uint64(0)
        self._check(self._lib.msc_fetch_matches(self._ctx, C.byref(out), C.byref(n)))
        try:
            if n.value == 0:
                return np.zeros(0, dtype=MATCH_DTYPE)
            buf = C.string_at(out, n.value * C.sizeof(_capi.msc_match))
            return np.frombuffer(buf, dtype=MATCH_DTYPE).copy()
        finally:
            self._lib.msc_free(out)

    def nonmatch_ids(self) -> np.ndarray:
        """Ids of the reads without a confirmed match, ascending (device-side, msc_fetch_nonmatch)."""
        n = C.c_uint64(0)
        self._check(self._lib.msc_fetch_nonmatch(self._ctx, None, 0, C.byref(n)))
        ids = np.empty(max(1, int(n.value)), dtype=np.uint32)
        self._check(self._lib.msc_fetch_nonmatch(self._ctx, ids.ctypes.data, len(ids), C.byref(n)))
        return ids[: int(n.value)]

    def fetch_into(self, dst_ptr: int, capacity: int) -> int:
        """Copy the matches into a caller-owned (e.g. pinned) buffer of `capacity` 16-byte records."""
        n = C.c_uint64(0)
        self._check(self._lib.msc_fetch_matches_into(self._ctx, dst_ptr, capacity, C.byref(n)))
        return int(n.value)

    def dump_keys(self) -> np.ndarray:
        out = C.POINTER(_capi.msc_key_rec)()
        n = C.c_uint64(0)
        self._check(self._lib.msc_dump_keys(self._ctx, C.byref(out), C.byref(n)))
        try:
            buf = C.string_at(out, n.value * C.sizeof(_capi.msc_key_rec)) if n.value else b""
            return np.frombuffer(buf, dtype=KEY_DTYPE).copy()
        finally:
            self._lib.msc_free(out)

    def dump_candidates(self) -> np.ndarray:
        out = C.POINTER(_capi.msc_cand_rec)()
        n = C.c_uint64(0)
        self._check(self._lib.msc_dump_candidates(self._ctx, C.byref(out), C.byref(n)))
        try:
            buf = C.string_at(out, n.value * C.sizeof(_capi.msc_cand_rec)) if n.value else b""
            return np.frombuffer(buf, dtype=CAND_DTYPE).copy()
        finally:
            self._lib.msc_free(out)

    def set_stage_timing(self, on: bool):
        """Per-stage timers on/off (one event record per stage boundary); the scan kernel is always timed."""
        self._check(self._lib.msc_set_stage_timing(self._ctx, 1 if on else 0))

    def stats(self) -> dict:
        st = _capi.msc_stats()
        self._check(self._lib.msc_get_stats(self._ctx, C.byref(st)))
        return st.as_dict()

    def reset_stats(self):
        self._lib.msc_reset_stats(self._ctx)


# ---------------------------------------------------------------------------------------------
# Inputs larger than one context call: read batches x target ranges.
#
# The reference streams inputs of any size (cmd/muscato_screen/main.go:408-480 reads the target
# file line by line; cmd/muscato_confirm/main.go:98-148 streams the sorted record files).  One
# msc_set_reads call takes at most 2^30 (read, window) items and one msc_set_targets call fewer than
# 2^32 - 4096 bases (include/muscato_b200.h), so larger inputs are cut into tiles.  Reads are
# independent of each other -- every rule of the path, MMTol included, is per read -- so read batches
# need no exchange at all; target ranges of one read batch share the per-read minimum of the MMTol
# rule (cmd/muscato_combine_windows/main.go:36-60): every tile is combined against its LOCAL minimum
# (a superset of what survives globally, because the local minimum is >= the global one) and the
# union is filtered once more against the minimum over all ranges.
# ---------------------------------------------------------------------------------------------
MAX_ITEMS_PER_READ_SET = 1 << 30
MAX_BASES_PER_TARGET_SET = (1 << 32) - 8192


def split_read_batches(n_reads: int, n_windows: int, max_items: int = MAX_ITEMS_PER_READ_SET):
    """[lo, hi) read-index ranges with (hi - lo) * n_windows <= max_items, sizes as equal as possible."""
    if n_reads <= 0:
        return [(0, 0)]
    per = max(1, max_items // max(1, n_windows))
    nb = (n_reads + per - 1) // per
    return [(n_reads * i // nb, n_reads * (i + 1) // nb) for i in range(nb)]


def split_target_ranges(offs: np.ndarray, max_bases: int = MAX_BASES_PER_TARGET_SET):
    """[lo, hi) target-index ranges of whole targets with at most max_bases bases each (greedy)."""
    offs = np.asarray(offs)
    G = len(offs) - 1
    if G <= 0:
        return [(0, 0)]
    out = []
    lo = 0
    while lo < G:
        limit = int(offs[lo]) + max_bases
        hi = int(np.searchsorted(offs, limit, side="right")) - 1
        if hi <= lo:
            raise ValueError(f"target {lo} alone exceeds {max_bases} bases")
        hi = min(hi, G)
        out.append((lo, hi))
        lo = hi
    return out


def filter_by_global_best(parts, mmtol: int) -> np.ndarray:
    """Union of per-range match arrays of ONE read batch (each already combined against its local
    per-read minimum) -> the matches with nx <= global per-read minimum + MMTol, ordered by
    (read, gene, pos)."""
    parts = [p for p in parts if len(p)]
    if not parts:
        return np.zeros(0, dtype=MATCH_DTYPE)
    if len(parts) == 1:
        return parts[0]
    m = np.concatenate(parts)
    rid = m["read_id"].astype(np.int64)
    order = np.lexsort((m["pos"], m["gene_id"], rid))
    m = m[order]
    rid = rid[order]
    first = np.ones(len(m), dtype=bool)
    first[1:] = rid[1:] != rid[:-1]
    seg = np.cumsum(first) - 1
    best = np.minimum.reduceat(m["nx"].astype(np.int64), np.nonzero(first)[0])
    return m[m["nx"].astype(np.int64) <= best[seg] + int(mmtol)]


def run_tiled(hp: "HotPath", reads, targets, max_items: int = MAX_ITEMS_PER_READ_SET,
              max_bases: int = MAX_BASES_PER_TARGET_SET, on_tile=None) -> np.ndarray:
    """The whole hot path over inputs of any size on one context: loops over read batches and target
    ranges (see above), returns the matches with GLOBAL read and gene indices ordered by
    (read, gene, pos).  reads / targets = (ascii uint8, offs uint64[n+1]).  MaxMatches truncation
    is exact within a tile; a (window, k-mer) group that exceeds MaxMatches only through several
    target ranges together is reported by raising MuscatoError (use one target range, or the sharded
    protocol of dist.py, for such inputs)."""
    ra, ro = HotPath._as_arrays(reads)
    ta, to = HotPath._as_arrays(targets)
    nwin = len(hp.cfg.Windows)
    batches = split_read_batches(len(ro) - 1, nwin, max_items)
    ranges = split_target_ranges(to, max_bases)
    if len(ranges) > 1:
        hp.set_shards(len(ranges))       # flag key groups above MaxMatches / n_ranges instead of truncating per range
    out = []
    try:
        for (r0, r1) in batches:
            a0, a1 = int(ro[r0]), int(ro[r1])
            hp.set_reads((ra[a0:a1], ro[r0:r1 + 1] - ro[r0]))
            parts = []
            for (g0, g1) in ranges:
                b0, b1 = int(to[g0]), int(to[g1])
                hp.set_targets((ta[b0:b1], to[g0:g1 + 1] - to[g0]))
                hp.run()
                if len(ranges) > 1 and hp.shard_overflow():
                    raise MuscatoError(_capi.MSC_ERR_STATE, "a key group may exceed MaxMatches across target ranges: "
                                       "run it as one range or through dist.sharded_matches")
                m = hp.fetch()
                if on_tile is not None:
                    on_tile((r0, r1), (g0, g1), hp.stats())
                if len(m):
                    m["gene_id"] += np.uint32(g0)
                    m["read_id"] += np.uint32(r0)
                parts.append(m)
            out.append(filter_by_global_best(parts, hp.cfg.MMTol))
    finally:
        if len(ranges) > 1:
            hp.set_shards(1)
    out = [p for p in out if len(p)]
    return np.concatenate(out) if out else np.zeros(0, dtype=MATCH_DTYPE)


# ---------------------------------------------------------------------------------------------
# Read parts on ONE GPU: upload / compute overlap.
#
# Every rule of the path is per read (window validity, confirm, MMTol, the result order), so a read
# set can be screened as independent parts, each with its own key table.  With the parts in separate
# contexts (separate streams) the upload of part i+1 runs under the scan + confirm + combine of part
# i: the end-to-end step of BASELINE configs[2] is PCIe time + the LAST part's pipeline instead of
# PCIe time + the whole pipeline.  Only the first context uploads the targets; the others copy the
# packed form device to device (msc_set_targets_from).  MaxMatches bounds a (window, k-mer) group
# over ALL its reads: the parts run in sharded mode (groups above MaxMatches / parts are flagged)
# and a flagged step raises, so that the caller falls back to one context.
# ---------------------------------------------------------------------------------------------
class ReadParts:
    def __init__(self, cfg: Config, parts: int = 2, device: int = 0, keep_ascii: bool = True):
        self.parts = int(parts)
        self.ctx = [HotPath(cfg, device=device, keep_ascii=keep_ascii) for _ in range(self.parts)]
        if self.parts > 1:
            for hp in self.ctx:
                hp.set_shards(self.parts)

    def close(self):
        for hp in self.ctx:
            hp.close()
        self.ctx = []

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    def step(self, reads_ptrs, targets_ptr, out_ptrs):
        """reads_ptrs[i] = (ascii_ptr, offs_ptr, n_reads) of part i (offsets relative to the part), targets_ptr =
        (ascii_ptr, offs_ptr, n_targets), out_ptrs[i] = (dst_ptr, capacity in 16-byte records): host (pinned)
        buffers.  Returns the number of matches of every part; read ids are relative to the part."""
        res = [0] * self.parts
        err = []

        def finish(i):
            try:
                hp = self.ctx[i]
                hp.run_stages(0, 1 | 2 | 4)
                if self.parts > 1 and hp.shard_overflow():
                    raise MuscatoError(_capi.MSC_ERR_STATE, "a key group may exceed MaxMatches across read parts: run the "
                                       "reads through one context")
                res[i] = hp.fetch_into(*out_ptrs[i])
            except Exception as e:  # surfaced by the caller's thread
                err.append(e)

        threads = []
        for i, hp in enumerate(self.ctx):
            if i == 0:
                hp.set_reads_ptr(*reads_ptrs[0])
                hp.set_targets_ptr(*targets_ptr)   # the target upload runs under part 0's table build
            else:
                hp.set_targets_from(self.ctx[0])
                hp.set_reads_ptr(*reads_ptrs[i])   # ... and this upload under part i-1's scan + confirm + combine
            t = threading.Thread(target=finish, args=(i,))
            t.start()
            threads.append(t)
        for t in threads:
            t.join()
        if err:
            raise err[0]
        return res

    def stats(self):
        return [hp.stats() for hp in self.ctx]

    def reset_stats(self):
        for hp in self.ctx:
            hp.reset_stats()
