"""CPU-only checks: host-side mirrors, the .sz codec, and that the C-ABI library loads and
exports every symbol include/muscato_b200.h declares (no compute calls without a GPU)."""
import ctypes
import os
import re

import numpy as np
import pytest

from muscato_b200 import _capi, build, formats, gendat, sz
from muscato_b200.config import Config
from tests import helpers


def test_nmiss_float64_truncation_table():
    """int((1-PMatch)*float64(L)), cmd/muscato_confirm/main.go:198."""
    table = {
        0.99: [0, 0, 1, 1, 2, 3], 0.98: [1, 1, 2, 3, 4, 6], 0.97: [1, 2, 3, 4, 6, 9],
        0.96: [2, 3, 4, 6, 8, 12], 0.95: [2, 3, 5, 7, 10, 15], 0.93: [3, 5, 6, 10, 13, 20],
        0.9: [4, 7, 9, 14, 19, 29],
    }
    for p, want in table.items():
        cfg = Config(Windows=[0], WindowWidth=4, MaxReadLength=300, PMatch=p)
        assert [cfg.nmiss(L) for L in (50, 75, 100, 150, 200, 300)] == want


def test_config_defaults_and_mandatory_fields():
    with pytest.raises(ValueError):
        Config(WindowWidth=4, MaxReadLength=10).apply_defaults()
    c = Config(Windows=[0, 5], WindowWidth=4, MaxReadLength=10).apply_defaults()
    assert (c.BloomSize, c.NumHash, c.PMatch, c.MaxMatches, c.MaxConfirmProcs, c.MatchMode, c.SortPar, c.SortMem) == \
        (4000000000, 20, 1.0, 1000000, 3, "best", 8, "50%")
    assert (c.MinDinuc, c.MMTol, c.MinReadLength) == (0, 0, 0)
    m = c.to_msc(device=0)
    assert m.n_windows == 2 and list(m.windows)[:2] == [0, 5] and m.window_width == 4 and m.match_mode == _capi.MSC_MATCH_BEST


def test_config_from_fixture_json():
    c = Config.from_json(os.path.join(helpers.GOLDEN, "muscato", "00", "config.json"))
    assert c.Windows == [0, 5] and c.WindowWidth == 4 and c.MMTol == 1 and c.MaxMatches == 1000


def test_sz_reads_reference_fixture_and_round_trips():
    d = sz.read_file(os.path.join(helpers.GOLDEN, "prep_targets", "06", "genes.txt.sz"))
    assert d == b"gene1\tATACGATCTACGATCA\ngene2\tTTAATTAATTAA\ngene3\tATTAGGCC\n"
    rng = np.random.default_rng(0)
    blob = helpers.random_dna(rng, 200000) + b"\n"
    assert sz.decompress(sz.compress(blob)) == blob
    assert sz.compress(b"")[:10] == b"\xff\x06\x00\x00sNaPpY"
    assert sz.crc32c(b"123456789") == 0xE3069283
    bad = bytearray(sz.compress(b"hello world\n"))
    bad[-1] ^= 1
    with pytest.raises(ValueError):
        sz.decompress(bytes(bad))


def test_prep_reads_mirror_matches_oracle(tmp_path, oracle_bin):
    rng = np.random.default_rng(3)
    raw = [helpers.random_dna(rng, int(rng.integers(5, 40)), b"ACGTN") for _ in range(200)]
    raw += raw[:30]
    names = [b"@r%d extra" % i for i in range(len(raw))]
    genes = [helpers.random_dna(rng, 80) for _ in range(3)]
    fq, gs, gi = helpers.write_case(str(tmp_path), raw, names, genes)
    cfgd = dict(Windows=[0], WindowWidth=4, MaxReadLength=30, MinReadLength=8, PMatch=1.0, BloomSize=100000, NumHash=4)
    out = helpers.oracle_pipeline(str(tmp_path), fq, gs, gi, cfgd, sub="upstream")
    got = formats.prep_reads_uniqify(helpers.read_bytes(fq), 8, 30)
    assert got == formats.load_reads_sorted(out["reads_sorted"])


def test_nonmatch_name_rule():
    assert formats.nonmatch_name("data/x/result.txt") == "data/x/result.nonmatch.txt.fastq"
    assert formats.nonmatch_name("results") == "nonmatch.results.fastq"


def test_gendat_plants_reads_like_reference():
    syn = gendat.generate(500, 50, 40, 300, seed=7)
    reads = set(syn.reads_list())
    tg = syn.targets_list()
    planted = 0
    for i in range(20):  # i < NumGene/2: read i%10 at offset i%10 (cmd/muscato_gendat/main.go:122-125)
        j = i % 10
        planted += tg[i][j:j + 50] in reads
    assert planted == 20
    syn_r = gendat.generate(500, 50, 40, 300, seed=7, rev=True)
    assert syn_r.n_targets == 80
    t0, t1 = syn_r.targets_list()[:2]
    comp = {65: 84, 84: 65, 71: 67, 67: 71}
    assert bytes(comp[c] for c in reversed(t0)) == t1


def _header_functions():
    hdr = open(os.path.join(helpers.ROOT, "include", "muscato_b200.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    return sorted(set(re.findall(r"\b(msc_[a-z_]+)\s*\(", hdr)))


def test_library_builds_loads_and_exports_every_declared_symbol():
    build.build()
    lib = ctypes.CDLL(build.LIB_PATH)
    declared = _header_functions()
    assert len(declared) >= 15
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in include/muscato_b200.h but not exported"
    assert sorted(_capi.EXPORTED_SYMBOLS) == declared
    lib.msc_version.restype = ctypes.c_char_p
    assert b"sm_100a" in lib.msc_version()


def test_struct_layouts_match_the_library():
    lib = _capi.load()
    for which, st in enumerate([_capi.msc_config, _capi.msc_match, _capi.msc_stats, _capi.msc_key_rec,
                                _capi.msc_cand_rec]):
        assert lib.msc_struct_size(which) == ctypes.sizeof(st), st.__name__
    assert ctypes.sizeof(_capi.msc_match) == 16


def test_no_gpu_means_loud_failure_not_fallback():
    """Without a usable device msc_create must fail with a message (no CPU path exists)."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from muscato_b200.engine import HotPath, MuscatoError
    with pytest.raises(MuscatoError) as ei:
        HotPath(Config(Windows=[0], WindowWidth=4, MaxReadLength=10).apply_defaults())
    assert "CUDA" in str(ei.value) or "device" in str(ei.value)


def test_product_never_touches_the_oracle():
    """Only tests/, __graft_entry__.smoke() and bench.py's baseline legs may reference oracle/."""
    pkg = os.path.join(helpers.ROOT, "muscato_b200")
    for root, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cc", ".h", ".cpp")):
                txt = open(os.path.join(root, f), errors="replace").read()
                assert "oracle/" not in txt.replace("write_oracle_inputs", "") or f == "gendat.py", f
                assert "muscato_oracle" not in txt, f


def test_cpp_sz_codec_matches_python_and_reference_fixture(tmp_path):
    """The C++ host executable's .sz codec (sztool -d / -c equivalents) against the reference's own
    compressed fixture and the Python codec."""
    import subprocess
    build.build()
    exe = build.EXE_PATH
    src = os.path.join(helpers.GOLDEN, "prep_targets", "06", "genes.txt.sz")
    out = subprocess.run([exe, "--sz-cat", src], capture_output=True, check=True).stdout
    assert out == sz.read_file(src)
    rng = np.random.default_rng(4)
    blob = b"\n".join(helpers.random_dna(rng, int(rng.integers(1, 300))) for _ in range(2000)) + b"\n"
    plain, packed = tmp_path / "x.txt", tmp_path / "x.txt.sz"
    plain.write_bytes(blob)
    subprocess.run([exe, "--sz-pack", str(plain), str(packed)], check=True)
    assert sz.read_file(str(packed)) == blob                       # Python reads what C++ wrote
    sz.write_file(str(tmp_path / "y.txt.sz"), blob)
    out = subprocess.run([exe, "--sz-cat", str(tmp_path / "y.txt.sz")], capture_output=True, check=True).stdout
    assert out == blob                                             # C++ reads what Python wrote


def test_cpp_executable_fails_loudly_without_gpu(tmp_path):
    import json
    import subprocess
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    build.build()
    src = os.path.join(helpers.GOLDEN, "muscato", "00")
    cfgd = json.load(open(os.path.join(src, "config.json")))
    seq, ids = str(tmp_path / "g.txt"), str(tmp_path / "gi.txt")
    helpers.oracle_prep_targets(os.path.join(src, "genes.txt"), seq, ids)
    cfgd.update(ReadFileName=os.path.join(src, "reads.fastq"), GeneFileName=seq, GeneIdFileName=ids,
                ResultsFileName=str(tmp_path / "result.txt"), TempDir=str(tmp_path))
    json.dump(cfgd, open(tmp_path / "config.json", "w"))
    r = subprocess.run([build.EXE_PATH, str(tmp_path / "config.json"), "--from-fastq"], capture_output=True, text=True)
    assert r.returncode != 0 and "CUDA" in r.stderr
    assert not os.path.exists(tmp_path / "result.txt")            # nothing is produced by a fallback


def test_host_half_of_device_prep_reads():
    """formats.uniqify_from_groups (the host half of msc_prep_reads) fed with the grouping a
    bytewise sort produces must reproduce prepReads (prep_reads | sort | uniqify) exactly."""
    import numpy as np
    from muscato_b200 import formats
    rng = np.random.default_rng(5)
    recs = []
    pool = []
    for i in range(400):
        if pool and rng.random() < 0.4:
            s = pool[int(rng.integers(len(pool)))]
            if rng.random() < 0.3:
                s = s[: max(1, len(s) // 2)]
        else:
            s = helpers.random_dna(rng, int(rng.integers(5, 40)), b"ACGTN")
        pool.append(s)
        recs.append((b"@n%d\tx y" % int(rng.integers(50)), s))
    fq = b"".join(n + b"\n" + s + b"\n+\n" + b"!" * len(s) + b"\n" for n, s in recs)
    min_len, max_len = 8, 30
    want = formats.prep_reads_uniqify(fq, min_len, max_len)
    names, raw = formats.parse_fastq(fq)
    # what the device returns: kept reads sorted by (X-substituted, truncated) sequence, group starts
    def key(s):
        return bytes(c if c in b"ATCG" else 0x58 for c in s[:max_len])
    kept = [i for i, s in enumerate(raw) if len(s) >= min_len]
    perm = sorted(kept, key=lambda i: key(raw[i]))
    gs = [0]
    for j in range(1, len(perm)):
        if key(raw[perm[j]]) != key(raw[perm[j - 1]]):
            gs.append(j)
    gs.append(len(perm))
    got = formats.uniqify_from_groups(names, raw, np.array(perm, dtype=np.uint32), np.array(gs, dtype=np.uint32), max_len)
    assert got == want
    # non-match fastq from ids == from matches
    m = np.zeros(3, dtype=[("read_id", "u4"), ("gene_id", "u4"), ("pos", "u4"), ("nx", "u4")])
    m["read_id"] = [0, 2, 2]
    ids = np.setdiff1d(np.arange(len(want[0])), np.unique(m["read_id"]))
    assert formats.nonmatch_fastq_from_ids(ids, *want) == formats.nonmatch_fastq(m, *want)


def test_cpp_sz_decoder_parallel(tmp_path):
    """csrc/host/szio.hpp: the C++ framed-snappy reader (chunks decoded by several host threads)
    against the Python codec and the reference's own compressed fixtures."""
    import subprocess
    import numpy as np
    from muscato_b200 import sz
    build.build()
    rng = np.random.default_rng(1)
    # multi-chunk file written by the Python writer (stored chunks), ~3 MB of text
    txt = b"\n".join(helpers.random_dna(rng, int(rng.integers(20, 3000)), b"ACGTX") for _ in range(2000)) + b"\n"
    p = str(tmp_path / "big.txt.sz")
    sz.write_file(p, txt)
    for threads in ("1", "4", "0"):
        r = subprocess.run([build.EXE_PATH, "--sz-cat", p, threads], capture_output=True)
        assert r.returncode == 0, r.stderr
        assert r.stdout == txt
    # Snappy-compressed chunks (type 0x00) as golang/snappy wrote them
    for case in ("06", "07"):
        f = os.path.join(helpers.GOLDEN, "prep_targets", case, "genes.txt.sz")
        r = subprocess.run([build.EXE_PATH, "--sz-cat", f, "3"], capture_output=True)
        assert r.returncode == 0, r.stderr
        assert r.stdout == sz.read_file(f)
    # a corrupted payload byte must be caught by the chunk CRC
    raw = bytearray(open(p, "rb").read())
    raw[len(raw) // 2] ^= 0x20
    q = str(tmp_path / "bad.txt.sz")
    open(q, "wb").write(bytes(raw))
    r = subprocess.run([build.EXE_PATH, "--sz-cat", q, "4"], capture_output=True)
    assert r.returncode != 0
