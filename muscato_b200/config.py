"""Mirror of utils.Config (utils/config.go:10-101) and of the driver's flag handling
(handleArgs / checkArgs, cmd/muscato/main.go:708-904) for the fields the hot path owns."""
from __future__ import annotations

import dataclasses
import json
from typing import List

from . import _capi


@dataclasses.dataclass
class Config:
    ReadFileName: str = ""
    GeneFileName: str = ""
    GeneIdFileName: str = ""
    ResultsFileName: str = ""
    Windows: List[int] = dataclasses.field(default_factory=list)
    WindowWidth: int = 0
    BloomSize: int = 0
    NumHash: int = 0
    PMatch: float = 0.0
    MinDinuc: int = 0
    TempDir: str = ""
    LogDir: str = ""
    MinReadLength: int = 0
    MaxReadLength: int = 0
    MaxMatches: int = 0
    MaxConfirmProcs: int = 0
    MMTol: int = 0
    MatchMode: str = ""
    SortPar: int = 0
    SortTemp: str = ""
    SortMem: str = ""
    NoCleanTemp: bool = False
    CPUProfile: bool = False

    @classmethod
    def from_json(cls, path: str) -> "Config":
        """utils.ReadConfig (utils/config.go:103-117): unknown keys are ignored."""
        with open(path, "rb") as f:
            raw = json.load(f)
        names = {f.name for f in dataclasses.fields(cls)}
        return cls(**{k: v for k, v in raw.items() if k in names and v is not None})

    def to_json(self, path: str) -> None:
        with open(path, "w") as f:
            json.dump(dataclasses.asdict(self), f)

    def apply_defaults(self) -> "Config":
        """checkArgs defaults (cmd/muscato/main.go:859-903); mandatory fields raise ValueError."""
        if not self.Windows:
            raise ValueError("Windows not provided")
        if self.WindowWidth == 0:
            raise ValueError("WindowWidth not provided")
        if self.MaxReadLength == 0:
            raise ValueError("MaxReadLength not provided")
        if self.BloomSize == 0:
            self.BloomSize = 4 * 1000 * 1000 * 1000
        if self.NumHash == 0:
            self.NumHash = 20
        if self.PMatch == 0:
            self.PMatch = 1.0
        if self.MaxMatches == 0:
            self.MaxMatches = 1000 * 1000
        if self.MaxConfirmProcs == 0:
            self.MaxConfirmProcs = 3
        if self.MatchMode == "":
            self.MatchMode = "best"
        if self.SortPar == 0:
            self.SortPar = 8
        if self.SortMem == "":
            self.SortMem = "50%"
        return self

    def nmiss(self, read_len: int) -> int:
        """int((1-PMatch)*float64(L)) -- cmd/muscato_confirm/main.go:198 (IEEE double, truncation)."""
        return int((1 - self.PMatch) * float(read_len))

    def to_msc(self, device: int = 0, keep_ascii: bool = False, bloom_bits_per_key: int = 0) -> "_capi.msc_config":
        c = _capi.msc_config()
        if len(self.Windows) > _capi.MSC_MAX_WINDOWS:
            raise ValueError("at most 32 windows are supported")
        c.n_windows = len(self.Windows)
        for i, w in enumerate(self.Windows):
            c.windows[i] = int(w)
        c.window_width = int(self.WindowWidth)
        c.max_read_length = int(self.MaxReadLength)
        c.pmatch = float(self.PMatch if self.PMatch != 0 else 1.0)
        c.min_dinuc = int(self.MinDinuc)
        c.mmtol = int(self.MMTol)
        c.max_matches = int(self.MaxMatches if self.MaxMatches else 1000 * 1000)
        mode = self.MatchMode or "best"
        if mode not in ("first", "best"):
            raise ValueError("MatchMode must be 'first' or 'best'")
        c.match_mode = _capi.MSC_MATCH_FIRST if mode == "first" else _capi.MSC_MATCH_BEST
        c.device = int(device)
        c.bloom_bits_per_key = int(bloom_bits_per_key)
        c.keep_ascii = 1 if keep_ascii else 0
        return c
