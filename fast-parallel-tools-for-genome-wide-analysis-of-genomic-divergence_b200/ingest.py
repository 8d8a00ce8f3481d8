"""VCF / GTrack ingest: text files -> per-chromosome genotype arrays in the layouts the scans take.

Reference (paths relative to /root/reference/):
  tools/VCFConvert.py:5-89            VCF -> GTrack text, one population at a time (`convertToGtrackFile`, `addHeader`)
  tools/ConvertVCFToGtrackTool.py:128-152   population from a file (one name per line) or a comma-separated box
  statistics/fisher/testFisher.c:193-227    GTrack points read back as (position, value) arrays
  statistics/FisherExactScoreStat.py:33-40  what the Statistic sees per chromosome: starts / vals of each population,
                                            position-major, individual-minor

Two routes to the same arrays:
  * `vcf_to_gtrack` + `read_gtrack`  — the reference's own two-step route (text in the middle), kept for drop-in use of
    existing GTrack files and for parity tests;
  * `read_vcf`                       — straight from VCF text to compact int8 codes, parsed by the native scanner in
    libfpt_b200.so (csrc/fpt_ingest.h); 8x less memory than the float64 route and no text round trip.
"""
import ctypes as C
from collections import OrderedDict

import numpy as np

from . import _lib

MISSING_F64 = -10000.0
MISSING_I8 = -128

# tools/VCFConvert.py:8-17
GT_VALUES = {"./.": -10000, ".|.": -10000, "1/0": 0, "0/1": 0, "1|0": 0, "0|1": 0, "0/0": 3, "0|0": 3, "1/1": -3, "1|1": -3}


class Population:
    """One population on one chromosome in the reference layout: `starts` int32 [nsnp*size], `vals` float64 [nsnp*size]
    (what TrackView.startsAsNumpyArray / valsAsNumpyArray hand to `_compute`)."""

    def __init__(self, starts, vals):
        self.starts = np.ascontiguousarray(starts, dtype=np.int32)
        self.vals = np.ascontiguousarray(vals, dtype=np.float64)

    def startsAsNumpyArray(self):
        return self.starts

    def valsAsNumpyArray(self):
        return self.vals

    @property
    def size(self):
        """individuals per SNP: the run of equal leading positions, as the C code finds it (comparative.c:25-34)"""
        if self.starts.size == 0:
            return 0
        n = int(np.argmax(self.starts != self.starts[0]))
        return n if n > 0 else int(self.starts.size)


class CompactChromosome:
    """Both populations on one chromosome in the compact layout: pos int32 [nsnp], acodes int8 [nsnp, asize],
    bcodes int8 [nsnp, bsize]."""

    def __init__(self, pos, acodes, bcodes):
        self.pos = np.ascontiguousarray(pos, dtype=np.int32)
        self.acodes = np.ascontiguousarray(acodes, dtype=np.int8)
        self.bcodes = np.ascontiguousarray(bcodes, dtype=np.int8)

    @property
    def nsnp(self):
        return int(self.pos.size)

    def reference_layout(self):
        """-> (Population A, Population B) as float64 / repeated positions"""
        out = []
        for codes in (self.acodes, self.bcodes):
            vals = codes.astype(np.float64)
            vals[codes == MISSING_I8] = MISSING_F64
            out.append(Population(np.repeat(self.pos, codes.shape[1]), vals.reshape(-1)))
        return tuple(out)


# ------------------------------------------------------------------------------------------------ populations
def read_population(text):
    """one individual per line (ConvertVCFToGtrackTool.py:131-136: only the newline is stripped)"""
    return [line.rstrip("\n") for line in text.splitlines(True)]


def parse_population_box(text):
    """comma-separated list, whitespace stripped (ConvertVCFToGtrackTool.py:138)"""
    return [x.strip() for x in text.split(",")]


# ------------------------------------------------------------------------------------------------ VCF -> GTrack text
def gtrack_header(genome):
    """tools/VCFConvert.py:49-53"""
    return ("##gtrack version: 1.0\n" "##track type: valued points\n" "##value type: number\n###seqid\t"
            + ("start\tvalue\tgenomeid\n####genome=%s\n" % genome))


def _vcf_columns(header_line, pop):
    header = header_line.split("\t")
    chromidx, posidx, formatidx = header.index("#CHROM"), header.index("POS"), header.index("FORMAT")
    found, cols, missing = [], [], []
    for name in pop:                                  # VCFConvert.py:55-68: unknown names are reported and dropped
        if name in header:
            found.append(name)
            cols.append(header.index(name))
        else:
            missing.append(name)
    return chromidx, posidx, formatidx, cols, found, missing


def vcf_to_gtrack(data, pop, genome="test", with_header=True):
    """The reference's conversion, string for string (`addHeader(genome) + convertToGtrackFile(data, pop, genome)`).
    Returns (text, individuals_not_found)."""
    lines = data.strip().split("\n")
    idx = 0
    while not lines[idx].startswith("#CHROM"):
        idx += 1
    chromidx, posidx, formatidx, cols, found, missing = _vcf_columns(lines[idx], list(pop))
    gtidx = lines[idx + 1].split("\t")[formatidx].split(":").index("GT")
    out = [gtrack_header(genome)] if with_header else []
    for line in lines[idx + 1:]:
        words = line.split("\t")
        chrom, pos = words[chromidx], words[posidx]
        for col, name in zip(cols, found):
            out.append("%s\t%s\t%s\t%s\n" % (chrom, pos, GT_VALUES[words[col].split(":")[gtidx]], name))
    return "".join(out), missing


# ------------------------------------------------------------------------------------------------ native scanners
def _as_bytes(text):
    return text if isinstance(text, (bytes, bytearray)) else text.encode("utf-8")


def _runs_to_groups(buf, runs, nruns, nrecords):
    """chromosome runs -> OrderedDict name -> list of (first, last) record ranges, in order of first appearance"""
    groups = OrderedDict()
    for k in range(nruns):
        name = buf[runs[k].name_off:runs[k].name_off + runs[k].name_len].decode("utf-8")
        first = runs[k].first_record
        last = runs[k + 1].first_record if k + 1 < nruns else nrecords
        groups.setdefault(name, []).append((first, last))
    return groups


def _take(arr, ranges):
    return arr[ranges[0][0]:ranges[0][1]] if len(ranges) == 1 else np.concatenate([arr[a:b] for a, b in ranges])


def _sort_by_pos(pos, *arrays):
    """the track preprocessor orders points by start; stable, so individuals keep their file order inside a SNP"""
    if pos.size > 1 and np.any(pos[1:] < pos[:-1]):
        order = np.argsort(pos, kind="stable")
        return (pos[order],) + tuple(a[order] for a in arrays)
    return (pos,) + arrays


def read_vcf(data, pop_a, pop_b):
    """VCF text (str or bytes) -> (OrderedDict chromosome -> CompactChromosome, info). `info` lists the individuals of
    each population that the header does not name (dropped, as VCFConvert.py:55-68 does) and the ones used."""
    lib = _lib.load()
    buf = _as_bytes(data)
    hoff, boff, nrec = C.c_int64(), C.c_int64(), C.c_int64()
    _lib.check(lib.fpt_vcf_scan(buf, len(buf), C.byref(hoff), C.byref(boff), C.byref(nrec)))
    hend = buf.find(b"\n", hoff.value)
    header_line = buf[hoff.value:hend if hend >= 0 else len(buf)].decode("utf-8").rstrip("\r")
    chromidx, posidx, formatidx, cols_a, found_a, miss_a = _vcf_columns(header_line, list(pop_a))
    _, _, _, cols_b, found_b, miss_b = _vcf_columns(header_line, list(pop_b))
    info = {"a": found_a, "b": found_b, "a_not_found": miss_a, "b_not_found": miss_b, "records": nrec.value}
    n, na, nb = nrec.value, len(cols_a), len(cols_b)
    cols = np.ascontiguousarray(cols_a + cols_b, dtype=np.int32)
    codes = np.empty((n, na + nb), dtype=np.int8)
    pos = np.empty(n, dtype=np.int32)
    max_runs = max(16, n)
    runs = (_lib.ChromRun * max_runs)()
    nruns = C.c_int64()
    if n > 0:
        _lib.check(lib.fpt_vcf_parse(buf, len(buf), boff.value, chromidx, posidx, formatidx, cols.ctypes.data, na + nb, n,
                                     codes.ctypes.data, pos.ctypes.data, C.cast(runs, C.c_void_p), max_runs, C.byref(nruns)))
    out = OrderedDict()
    for name, ranges in _runs_to_groups(buf, runs, nruns.value, n).items():
        p, c = _sort_by_pos(_take(pos, ranges), _take(codes, ranges))
        out[name] = CompactChromosome(p, c[:, :na], c[:, na:])
    return out, info


def read_gtrack(data):
    """GTrack valued-points text -> OrderedDict chromosome -> Population (reference layout). The column order comes from
    the `###` line when present (default: seqid start value genomeid, VCFConvert.py:52-53)."""
    lib = _lib.load()
    buf = _as_bytes(data)
    names = ["seqid", "start", "value", "genomeid"]
    for line in buf[:4096].split(b"\n"):
        if line.startswith(b"###") and not line.startswith(b"####"):
            names = line[3:].decode("utf-8").rstrip("\r").split("\t")
            break
    nrec = C.c_int64()
    _lib.check(lib.fpt_gtrack_scan(buf, len(buf), C.byref(nrec)))
    n = nrec.value
    pos = np.empty(n, dtype=np.int32)
    vals = np.empty(n, dtype=np.float64)
    max_runs = max(16, n)
    runs = (_lib.ChromRun * max_runs)()
    nruns = C.c_int64()
    if n > 0:
        _lib.check(lib.fpt_gtrack_parse(buf, len(buf), names.index("seqid"), names.index("start"), names.index("value"), n,
                                        pos.ctypes.data, vals.ctypes.data, C.cast(runs, C.c_void_p), max_runs, C.byref(nruns)))
    out = OrderedDict()
    for name, ranges in _runs_to_groups(buf, runs, nruns.value, n).items():
        p, v = _sort_by_pos(_take(pos, ranges), _take(vals, ranges))
        out[name] = Population(p, v)
    return out


def compact_codes(vals):
    """float64 reference values -> int8 codes (native loop; 3 / 0 / -3 kept, anything else -> -128)"""
    vals = np.ascontiguousarray(vals, dtype=np.float64)
    out = np.empty(vals.shape, dtype=np.int8)
    _lib.check(_lib.load().fpt_compact_codes(vals.ctypes.data, vals.size, out.ctypes.data))
    return out


def pair_populations(a, b):
    """two Populations of one chromosome -> CompactChromosome. Raises ValueError when they do not list the same SNP
    positions (the reference silently mis-pairs such input; the scans reject it with FPT_ERR_POSITIONS)."""
    asize, bsize = a.size, b.size
    if asize == 0 or bsize == 0 or a.starts.size % asize or b.starts.size % bsize:
        raise ValueError("population arrays are not a whole number of SNPs")
    pa, pb = a.starts[::asize], b.starts[::bsize]
    if pa.size != pb.size or not np.array_equal(pa, pb):
        raise ValueError("populations A and B list different SNP positions")
    if not (np.array_equal(np.repeat(pa, asize), a.starts) and np.array_equal(np.repeat(pb, bsize), b.starts)):
        raise ValueError("a SNP does not carry the same number of individuals as the first one")
    return CompactChromosome(pa, compact_codes(a.vals).reshape(-1, asize), compact_codes(b.vals).reshape(-1, bsize))
