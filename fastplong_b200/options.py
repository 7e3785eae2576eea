"""Host-side mirror of the reference's option handling for the hot path (src/main.cpp:27-250, src/options.h).

`Options` carries the CLI-level values with the reference's defaults and converts them to the flat
`fpl_options` POD + adapter list that the C ABI takes.  Flag names follow the reference CLI.
"""
import ctypes as C
from dataclasses import dataclass, field
from typing import List

from .abi import FplOptions, make_adapters


def num2qual(num: int) -> int:
    """util.h:260-268 — phred number to phred+33 char value, clamped like the reference."""
    if num > 127 - 33:
        num = 127 - 33
    if num < 0:
        num = 0
    return num + 33


def reverse_complement(s: str) -> str:
    """Sequence::reverseComplement (src/sequence.cpp:29-77): A/a->T, T/t->A, C/c->G, G/g->C, else N."""
    m = {"A": "T", "a": "T", "T": "A", "t": "A", "C": "G", "c": "G", "G": "C", "g": "C"}
    return "".join(m.get(ch, "N") for ch in reversed(s))


@dataclass
class Options:
    # adapter cutting (src/main.cpp:40-45, 129-144)
    start_adapter: str = "auto"
    end_adapter: str = "auto"
    adapter_fasta: List[str] = field(default_factory=list)  # sequences in std::map header order
    disable_adapter_trimming: bool = False
    distance_threshold: float = 0.25
    trimming_extension: int = 10
    # global trimming / quality cut (src/main.cpp:47-60, 146-177)
    trim_front: int = 0
    trim_tail: int = 0
    cut_front: bool = False
    cut_tail: bool = False
    cut_window_size: int = 4
    cut_mean_quality: int = 20
    cut_front_window_size: int = None
    cut_front_mean_quality: int = None
    cut_tail_window_size: int = None
    cut_tail_mean_quality: int = None
    # polyX (src/main.cpp:52-53)
    trim_poly_x: bool = False
    poly_x_min_len: int = 10
    # quality / length / complexity filters (src/main.cpp:76-90, 190-205)
    disable_quality_filtering: bool = False
    qualified_quality_phred: int = 15
    unqualified_percent_limit: int = 40
    mean_qual: int = 0
    n_percent_limit: int = 10
    n_base_limit: int = 1000000
    disable_length_filtering: bool = False
    length_required: int = 20
    length_limit: int = 0
    low_complexity_filter: bool = False
    complexity_threshold: int = 30
    # --mask / --break (src/main.cpp:62-70, 207-215)
    mask: bool = False
    mask_window_size: int = 50
    mask_mean_quality: int = 10
    break_reads: bool = False
    break_window_size: int = 100
    break_mean_quality: int = 10
    device: int = 0

    def resolve_adapters(self):
        """src/main.cpp:137-140: -s given and -e left at auto => -e = revcomp(-s)."""
        s, e = self.start_adapter, self.end_adapter
        if s != "auto" and e == "auto":
            e = reverse_complement(s)
        return s, e

    def to_abi(self):
        o = FplOptions()
        o.struct_size = C.sizeof(FplOptions)
        o.device = self.device
        o.trim_front, o.trim_tail = self.trim_front, self.trim_tail
        o.cut_front_enabled = int(self.cut_front)
        o.cut_tail_enabled = int(self.cut_tail)
        pick = lambda v, d: d if v is None else v  # noqa: E731
        o.cut_front_window = pick(self.cut_front_window_size, self.cut_window_size)
        o.cut_front_quality = pick(self.cut_front_mean_quality, self.cut_mean_quality)
        o.cut_tail_window = pick(self.cut_tail_window_size, self.cut_window_size)
        o.cut_tail_quality = pick(self.cut_tail_mean_quality, self.cut_mean_quality)
        o.polyx_enabled = int(self.trim_poly_x)
        o.polyx_min_len = self.poly_x_min_len
        o.adapter_enabled = int(not self.disable_adapter_trimming)
        o.trimming_extension = self.trimming_extension
        o.ed_max = self.distance_threshold
        o.qual_filter_enabled = int(not self.disable_quality_filtering)
        o.qualified_qual = num2qual(self.qualified_quality_phred)
        o.unqualified_percent_limit = self.unqualified_percent_limit
        o.avg_qual_req = self.mean_qual
        o.n_base_percent_limit = self.n_percent_limit
        o.n_base_limit = self.n_base_limit
        o.length_filter_enabled = int(not self.disable_length_filtering)
        o.length_required = self.length_required
        o.length_max = self.length_limit
        o.complexity_enabled = int(self.low_complexity_filter)
        o.complexity_threshold_pct = min(100, max(0, self.complexity_threshold))
        o.mask_enabled, o.mask_window, o.mask_quality = int(self.mask), self.mask_window_size, self.mask_mean_quality
        o.break_enabled, o.break_window, o.break_quality = (int(self.break_reads), self.break_window_size,
                                                            self.break_mean_quality)
        s, e = self.resolve_adapters()
        ad, keep = make_adapters(s, e, self.adapter_fasta)
        return o, ad, keep

    def adapter_list(self):
        s, e = self.resolve_adapters()
        return [s, e] + list(self.adapter_fasta)

    def cli_flags(self):
        """The equivalent reference command-line flags (for whole-binary runs of fastplong_ref / fastplong_gpu)."""
        f = []
        if self.start_adapter != "auto":
            f += ["-s", self.start_adapter]
        if self.end_adapter != "auto":
            f += ["-e", self.end_adapter]
        if self.disable_adapter_trimming:
            f += ["-A"]
        f += ["-d", repr(float(self.distance_threshold)), "--trimming_extension", str(self.trimming_extension)]
        if self.trim_front:
            f += ["-f", str(self.trim_front)]
        if self.trim_tail:
            f += ["-t", str(self.trim_tail)]
        if self.cut_front:
            f += ["--cut_front"]
        if self.cut_tail:
            f += ["--cut_tail"]
        f += ["-W", str(self.cut_window_size), "-M", str(self.cut_mean_quality)]
        for name, v in (("cut_front_window_size", self.cut_front_window_size),
                        ("cut_front_mean_quality", self.cut_front_mean_quality),
                        ("cut_tail_window_size", self.cut_tail_window_size),
                        ("cut_tail_mean_quality", self.cut_tail_mean_quality)):
            if v is not None:
                f += ["--" + name, str(v)]
        if self.trim_poly_x:
            f += ["-x"]
        f += ["--poly_x_min_len", str(self.poly_x_min_len)]
        if self.disable_quality_filtering:
            f += ["-Q"]
        f += ["-q", str(self.qualified_quality_phred), "-u", str(self.unqualified_percent_limit),
              "-m", str(self.mean_qual), "-n", str(self.n_percent_limit)]
        if self.n_base_limit != 1000000:
            f += ["--n_base_limit", str(self.n_base_limit)]
        if self.disable_length_filtering:
            f += ["-L"]
        f += ["-l", str(self.length_required), "--length_limit", str(self.length_limit)]
        if self.low_complexity_filter:
            f += ["-y", "-Y", str(self.complexity_threshold)]
        if self.mask:
            f += ["-N", "--mask_window_size", str(self.mask_window_size), "--mask_mean_quality", str(self.mask_mean_quality)]
        if self.break_reads:
            f += ["-b", "--break_window_size", str(self.break_window_size), "--break_mean_quality",
                  str(self.break_mean_quality)]
        return f
