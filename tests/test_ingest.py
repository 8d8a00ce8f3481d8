"""VCF / GTrack ingest against fixtures written by the reference's own converter
(tests/golden/make_golden.py: tools/VCFConvert.py run on tests/golden/small.vcf). CPU only: the scanners are host code
of libfpt_b200.so."""
import os

import numpy as np
import pytest

import fpt_b200.ingest as ingest
from fpt_b200._lib import FptError

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
POP_A = ["ind00", "ind02", "ind04", "ind07", "nobody"]
POP_B = ["ind08", "ind01", "ind03", "ind05"]


def _read(name):
    with open(os.path.join(GOLD, name)) as f:
        return f.read()


def test_vcf_to_gtrack_text_is_the_reference_converters():
    vcf = _read("small.vcf")
    for tag, pop in (("A", POP_A), ("B", POP_B)):
        text, missing = ingest.vcf_to_gtrack(vcf, pop, "test")
        assert text == _read("small_pop%s.gtrack" % tag)
        assert missing == (["nobody"] if tag == "A" else [])


def test_read_gtrack_layout():
    tracks = ingest.read_gtrack(_read("small_popA.gtrack"))
    assert list(tracks) == ["chrI", "chrII"]
    lines = [l.split("\t") for l in _read("small_popA.gtrack").split("\n") if l and l[0] != "#"]
    for chrom, tv in tracks.items():
        rows = [l for l in lines if l[0] == chrom]
        assert np.array_equal(tv.startsAsNumpyArray(), np.array([int(r[1]) for r in rows], dtype=np.int32))
        assert np.array_equal(tv.valsAsNumpyArray(), np.array([float(r[2]) for r in rows]))
        assert tv.size == 4 and tv.starts.dtype == np.int32 and tv.vals.dtype == np.float64
    assert tracks["chrI"].starts.size == 40 * 4 and tracks["chrII"].starts.size == 25 * 4


def test_read_vcf_equals_the_two_step_route():
    vcf = _read("small.vcf")
    direct, info = ingest.read_vcf(vcf, POP_A, POP_B)
    assert info["a_not_found"] == ["nobody"] and info["b_not_found"] == [] and info["records"] == 65
    assert info["a"] == POP_A[:4] and info["b"] == POP_B
    ta, tb = ingest.read_gtrack(_read("small_popA.gtrack")), ingest.read_gtrack(_read("small_popB.gtrack"))
    assert list(direct) == list(ta)
    for chrom, cc in direct.items():
        paired = ingest.pair_populations(ta[chrom], tb[chrom])
        assert np.array_equal(cc.pos, paired.pos)
        assert np.array_equal(cc.acodes, paired.acodes) and np.array_equal(cc.bcodes, paired.bcodes)
        a, b = cc.reference_layout()
        assert np.array_equal(a.vals, ta[chrom].vals) and np.array_equal(a.starts, ta[chrom].starts)
        assert np.array_equal(b.vals, tb[chrom].vals) and np.array_equal(b.starts, tb[chrom].starts)
        assert set(np.unique(cc.acodes)) <= {3, 0, -3, -128}


def test_vcf_edge_cases():
    head = "##x\n#CHROM\tPOS\tID\tREF\tALT\tQUAL\tFILTER\tINFO\tFORMAT\ts1\ts2\ts3\n"
    # CRLF line ends, trailing blank lines, chromosome re-appearing later, unsorted positions inside a chromosome
    body = ("c1\t30\t.\tA\tC\t.\t.\t.\tGT\t0/0\t1/1\t0/1\r\n" "c2\t5\t.\tA\tC\t.\t.\t.\tGT\t./.\t0|0\t1|1\r\n"
            "c1\t10\t.\tA\tC\t.\t.\t.\tGT\t1/1\t1/1\t0/0\r\n\r\n\n")
    out, info = ingest.read_vcf(head + body, ["s3", "s1"], ["s2"])
    assert list(out) == ["c1", "c2"] and info["records"] == 3
    assert out["c1"].pos.tolist() == [10, 30]
    assert out["c1"].acodes.tolist() == [[3, -3], [0, 3]] and out["c1"].bcodes.tolist() == [[-3], [-3]]
    assert out["c2"].acodes.tolist() == [[-3, -128]] and out["c2"].bcodes.tolist() == [[3]]
    # the GT slot is taken from the first record only (VCFConvert.py:31-33)
    body2 = "c1\t1\t.\tA\tC\t.\t.\t.\tDP:GT\t7:0/0\t7:1/1\t7:0/1\n" "c1\t2\t.\tA\tC\t.\t.\t.\tGT:DP\t9:0/1\t9:0/0\t9:1/1\n"
    out2, _ = ingest.read_vcf(head + body2, ["s1", "s2", "s3"], [])
    assert out2["c1"].acodes.tolist() == [[3, -3, 0], [0, 3, -3]] and out2["c1"].bcodes.shape == (2, 0)
    # header only
    out3, info3 = ingest.read_vcf(head, ["s1"], ["s2"])
    assert len(out3) == 0 and info3["records"] == 0


def test_vcf_errors():
    head = "#CHROM\tPOS\tID\tREF\tALT\tQUAL\tFILTER\tINFO\tFORMAT\ts1\ts2\n"
    with pytest.raises(FptError, match="not a diploid biallelic call"):       # KeyError in the reference
        ingest.read_vcf(head + "c1\t1\t.\tA\tC,G\t.\t.\t.\tGT\t0/2\t0/0\n", ["s1"], ["s2"])
    with pytest.raises(FptError, match="no #CHROM header"):
        ingest.read_vcf("c1\t1\t.\tA\tC\t.\t.\t.\tGT\t0/0\t0/0\n", ["s1"], ["s2"])
    with pytest.raises(FptError, match="POS"):
        ingest.read_vcf(head + "c1\tx\t.\tA\tC\t.\t.\t.\tGT\t0/0\t0/0\n", ["s1"], ["s2"])
    with pytest.raises(FptError, match="no GT slot"):
        ingest.read_vcf(head + "c1\t1\t.\tA\tC\t.\t.\t.\tDP\t3\t4\n", ["s1"], ["s2"])
    with pytest.raises(FptError, match="fewer than"):
        ingest.read_vcf(head + "c1\t1\t.\tA\tC\t.\t.\t.\tGT\t0/0\n", ["s1"], ["s2"])
    with pytest.raises(KeyError):
        ingest.vcf_to_gtrack(head + "c1\t1\t.\tA\tC,G\t.\t.\t.\tGT\t0/2\t0/0\n", ["s1"])


def test_gtrack_errors_and_column_order():
    text = "##gtrack version: 1.0\n###start\tseqid\tgenomeid\tvalue\n7\tc9\tx\t3\n7\tc9\ty\t-10000\n9\tc9\tx\t0\n9\tc9\ty\t-3\n"
    tv = ingest.read_gtrack(text)["c9"]
    assert tv.starts.tolist() == [7, 7, 9, 9] and tv.vals.tolist() == [3.0, -10000.0, 0.0, -3.0] and tv.size == 2
    with pytest.raises(FptError, match="not a number"):
        ingest.read_gtrack("c\t1\tabc\tx\n")
    with pytest.raises(FptError, match="start"):
        ingest.read_gtrack("c\t-4\t3\tx\n")
    assert len(ingest.read_gtrack("##gtrack version: 1.0\n")) == 0


def test_pair_populations_rejects_mismatched_positions():
    a = ingest.Population([1, 1, 5, 5], [3, 3, 0, 0])
    b = ingest.Population([1, 6], [3, -3])
    with pytest.raises(ValueError, match="different SNP positions"):
        ingest.pair_populations(a, b)
    with pytest.raises(ValueError, match="whole number"):
        ingest.pair_populations(ingest.Population([1, 1, 5], [3, 3, 0]), b)
    ok = ingest.pair_populations(a, ingest.Population([1, 5], [-10000, 7]))
    assert ok.acodes.tolist() == [[3, 3], [0, 0]] and ok.bcodes.tolist() == [[-128], [-128]]


def test_population_helpers():
    assert ingest.read_population("a\nb c\n d\n") == ["a", "b c", " d"]
    assert ingest.parse_population_box(" a, b ,c") == ["a", "b", "c"]
    assert ingest.gtrack_header("g").endswith("####genome=g\n") and ingest.gtrack_header("g").count("\n") == 5


def test_native_scanner_throughput_smoke():
    """100k records x 40 samples through the native scanner: seconds, not minutes"""
    import time
    rng = np.random.default_rng(0)
    names = ["s%d" % i for i in range(40)]
    gts = np.array(["0/0", "0/1", "1/1", "./."])
    calls = gts[rng.integers(0, 4, size=(100000, 40))]
    pos = np.sort(rng.choice(10**8, size=100000, replace=False))
    body = "\n".join("c1\t%d\t.\tA\tC\t.\t.\t.\tGT\t%s" % (p, "\t".join(row)) for p, row in zip(pos, calls))
    text = "#CHROM\tPOS\tID\tREF\tALT\tQUAL\tFILTER\tINFO\tFORMAT\t" + "\t".join(names) + "\n" + body + "\n"
    t = time.time()
    out, _ = ingest.read_vcf(text, names[:20], names[20:])
    dt = time.time() - t
    want = np.select([calls == "0/0", calls == "1/1", calls == "0/1"], [3, -3, 0], -128).astype(np.int8)
    assert np.array_equal(out["c1"].acodes, want[:, :20]) and np.array_equal(out["c1"].bcodes, want[:, 20:])
    assert np.array_equal(out["c1"].pos, pos.astype(np.int32))
    assert dt < 5.0
