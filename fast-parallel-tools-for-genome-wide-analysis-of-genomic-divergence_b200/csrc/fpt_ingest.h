/*
 * fpt_ingest.h — host-side text ingest: VCF and GTrack "valued points" files -> the compact genotype layout the
 * scans upload (one int8 code per individual and SNP, SNP-major; 3 / 0 / -3 / -128).
 *
 * Reference behaviour restated (paths relative to /root/reference/):
 *   tools/VCFConvert.py:5-17    GT string -> value table (0/0 -> 3, 1/1 -> -3, 0/1 1/0 -> 0, ./. -> -10000; '|' alike)
 *   tools/VCFConvert.py:22-33   the "#CHROM" line is the column header; the GT slot inside FORMAT is read from the FIRST
 *                               record only and applied to every record
 *   tools/VCFConvert.py:77-89   one output point per (record, selected individual), record-major
 *   tools/VCFConvert.py:49-53   GTrack header: five '#' lines, columns seqid start value genomeid
 *   statistics/fisher/testFisher.c:193-227  reading the points back: skip the header, strtol / strtod per line
 *
 * Included by fpt_api.cu (uses its fail()); plain C++ — nothing here touches the device.
 */
#ifndef FPT_INGEST_H
#define FPT_INGEST_H

#include <cerrno>
#include <cstdlib>
#include <cstring>

namespace fpt_ingest {

struct Cursor {
    const char *p, *end;
};

static inline const char *line_end(const char *p, const char *end) {
    const void *nl = memchr(p, '\n', (size_t)(end - p));
    return nl ? (const char *)nl : end;
}

/* the k-th tab-separated field of [p, e); false when the line has fewer fields */
static inline bool field(const char *p, const char *e, int k, const char **fs, const char **fe) {
    for (int i = 0; i < k; i++) {
        const void *t = memchr(p, '\t', (size_t)(e - p));
        if (!t) return false;
        p = (const char *)t + 1;
    }
    const void *t = memchr(p, '\t', (size_t)(e - p));
    *fs = p;
    *fe = t ? (const char *)t : e;
    return true;
}

/* the k-th ':'-separated slot of [p, e) */
static inline bool slot(const char *p, const char *e, int k, const char **fs, const char **fe) {
    for (int i = 0; i < k; i++) {
        const void *t = memchr(p, ':', (size_t)(e - p));
        if (!t) return false;
        p = (const char *)t + 1;
    }
    const void *t = memchr(p, ':', (size_t)(e - p));
    *fs = p;
    *fe = t ? (const char *)t : e;
    return true;
}

static inline int gt_code(const char *s, const char *e) {
    /* VCFConvert.py:8-17; anything else is a KeyError there, FPT_ERR_ARG here (returns 1 = unknown) */
    if (e - s != 3 || (s[1] != '/' && s[1] != '|')) return 1;
    char a = s[0], b = s[2];
    if (a == '.' && b == '.') return -128;
    if (a == '0' && b == '0') return 3;
    if (a == '1' && b == '1') return -3;
    if ((a == '0' && b == '1') || (a == '1' && b == '0')) return 0;
    return 1;
}

static inline bool parse_int(const char *s, const char *e, long long *out) {
    if (s == e) return false;
    bool neg = false;
    if (*s == '-' || *s == '+') { neg = *s == '-'; s++; }
    if (s == e) return false;
    long long v = 0;
    for (; s < e; s++) {
        if (*s < '0' || *s > '9') return false;
        v = v * 10 + (*s - '0');
        if (v > 4000000000LL) return false;
    }
    *out = neg ? -v : v;
    return true;
}

static inline const char *strip_cr(const char *s, const char *e) { return (e > s && e[-1] == '\r') ? e - 1 : e; }

}  // namespace fpt_ingest

/* Locate the "#CHROM" header line and count the records after it (every non-empty line; VCFConvert.py strips the text
   first, so trailing blank lines do not count). */
extern "C" int fpt_vcf_scan(const char *buf, size_t len, int64_t *header_off, int64_t *body_off, int64_t *nrecords) {
    using namespace fpt_ingest;
    if (!buf || !header_off || !body_off || !nrecords) return fail(FPT_ERR_ARG, "fpt_vcf_scan: null argument");
    const char *p = buf, *end = buf + len;
    *header_off = -1;
    while (p < end) {
        const char *e = line_end(p, end);
        if (e - p >= 6 && memcmp(p, "#CHROM", 6) == 0) { *header_off = p - buf; p = e < end ? e + 1 : end; break; }
        p = e < end ? e + 1 : end;
    }
    if (*header_off < 0) return fail(FPT_ERR_ARG, "VCF text has no #CHROM header line");
    *body_off = p - buf;
    int64_t n = 0;
    while (p < end) {
        const char *e = line_end(p, end);
        if (strip_cr(p, e) > p) n++;
        p = e < end ? e + 1 : end;
    }
    *nrecords = n;
    return FPT_OK;
}

/* Parse the records: pos[r], codes[r * nsamples + s] for the selected sample columns, and the chromosome as runs of
   consecutive records that share a CHROM string (name given as offset/length into buf). */
extern "C" int fpt_vcf_parse(const char *buf, size_t len, int64_t body_off, int chrom_col, int pos_col, int format_col,
                             const int32_t *sample_cols, int nsamples, int64_t nrecords, int8_t *codes, int32_t *pos,
                             fpt_chrom_run *runs, int64_t max_runs, int64_t *nruns) {
    using namespace fpt_ingest;
    if (!buf || !pos || !runs || !nruns || (nsamples > 0 && (!sample_cols || !codes)) || body_off < 0 || (size_t)body_off > len)
        return fail(FPT_ERR_ARG, "fpt_vcf_parse: bad argument");
    const char *p = buf + body_off, *end = buf + len;
    int64_t r = 0, nr = 0;
    int gtidx = -1;
    const char *run_s = nullptr;
    size_t run_n = 0;
    while (p < end && r < nrecords) {
        const char *e0 = line_end(p, end), *e = strip_cr(p, e0);
        if (e > p) {
            const char *fs, *fe;
            if (gtidx < 0) {                                      /* VCFConvert.py:31-33, 69-72: first record only */
                if (!field(p, e, format_col, &fs, &fe)) return fail(FPT_ERR_ARG, "VCF record %lld has no FORMAT column", (long long)r);
                int k = 0;
                const char *ss, *se;
                while (slot(fs, fe, k, &ss, &se)) {
                    if (se - ss == 2 && ss[0] == 'G' && ss[1] == 'T') { gtidx = k; break; }
                    k++;
                }
                if (gtidx < 0) return fail(FPT_ERR_ARG, "FORMAT of the first VCF record has no GT slot");
            }
            if (!field(p, e, chrom_col, &fs, &fe)) return fail(FPT_ERR_ARG, "VCF record %lld has no CHROM column", (long long)r);
            if (!run_s || (size_t)(fe - fs) != run_n || memcmp(fs, run_s, run_n) != 0) {
                if (nr >= max_runs) return fail(FPT_ERR_ARG, "more than %lld chromosome runs", (long long)max_runs);
                runs[nr].name_off = fs - buf;
                runs[nr].name_len = (int32_t)(fe - fs);
                runs[nr].reserved = 0;
                runs[nr].first_record = r;
                nr++;
                run_s = fs;
                run_n = (size_t)(fe - fs);
            }
            long long v;
            if (!field(p, e, pos_col, &fs, &fe) || !parse_int(fs, fe, &v) || v < 0 || v > 2147483647LL)
                return fail(FPT_ERR_ARG, "VCF record %lld: POS is not a non-negative 32-bit integer", (long long)r);
            pos[r] = (int32_t)v;
            /* sample columns are visited in file order when sorted, else by a fresh walk per column */
            int8_t *row = codes + r * (int64_t)nsamples;
            const char *cur = p;
            int curcol = 0;
            for (int s = 0; s < nsamples; s++) {
                int col = sample_cols[s];
                if (col < curcol) { cur = p; curcol = 0; }
                while (curcol < col) {
                    const void *t = memchr(cur, '\t', (size_t)(e - cur));
                    if (!t) return fail(FPT_ERR_ARG, "VCF record %lld has fewer than %d columns", (long long)r, col + 1);
                    cur = (const char *)t + 1;
                    curcol++;
                }
                const void *t = memchr(cur, '\t', (size_t)(e - cur));
                const char *ce = t ? (const char *)t : e;
                const char *ss, *se;
                if (!slot(cur, ce, gtidx, &ss, &se))
                    return fail(FPT_ERR_ARG, "VCF record %lld column %d has no slot %d", (long long)r, col, gtidx);
                int c = gt_code(ss, se);
                if (c == 1)
                    return fail(FPT_ERR_ARG, "VCF record %lld column %d: genotype '%.*s' is not a diploid biallelic call",
                                (long long)r, col, (int)(se - ss), ss);
                row[s] = (int8_t)c;
            }
            r++;
        }
        p = e0 < end ? e0 + 1 : end;
    }
    if (r != nrecords) return fail(FPT_ERR_ARG, "expected %lld VCF records, parsed %lld", (long long)nrecords, (long long)r);
    *nruns = nr;
    return FPT_OK;
}

/* GTrack points: count the data lines (not starting with '#', not empty) */
extern "C" int fpt_gtrack_scan(const char *buf, size_t len, int64_t *nrecords) {
    using namespace fpt_ingest;
    if (!buf || !nrecords) return fail(FPT_ERR_ARG, "fpt_gtrack_scan: null argument");
    const char *p = buf, *end = buf + len;
    int64_t n = 0;
    while (p < end) {
        const char *e = line_end(p, end);
        if (strip_cr(p, e) > p && *p != '#') n++;
        p = e < end ? e + 1 : end;
    }
    *nrecords = n;
    return FPT_OK;
}

/* GTrack points -> pos[r], vals[r], chromosome runs and genome-id runs (consecutive records with the same string) */
extern "C" int fpt_gtrack_parse(const char *buf, size_t len, int seqid_col, int start_col, int value_col, int64_t nrecords,
                                int32_t *pos, double *vals, fpt_chrom_run *runs, int64_t max_runs, int64_t *nruns) {
    using namespace fpt_ingest;
    if (!buf || !pos || !vals || !runs || !nruns) return fail(FPT_ERR_ARG, "fpt_gtrack_parse: null argument");
    const char *p = buf, *end = buf + len;
    int64_t r = 0, nr = 0;
    const char *run_s = nullptr;
    size_t run_n = 0;
    while (p < end && r < nrecords) {
        const char *e0 = line_end(p, end), *e = strip_cr(p, e0);
        if (e > p && *p != '#') {
            const char *fs, *fe;
            if (!field(p, e, seqid_col, &fs, &fe)) return fail(FPT_ERR_ARG, "GTrack record %lld has no seqid column", (long long)r);
            if (!run_s || (size_t)(fe - fs) != run_n || memcmp(fs, run_s, run_n) != 0) {
                if (nr >= max_runs) return fail(FPT_ERR_ARG, "more than %lld chromosome runs", (long long)max_runs);
                runs[nr].name_off = fs - buf;
                runs[nr].name_len = (int32_t)(fe - fs);
                runs[nr].reserved = 0;
                runs[nr].first_record = r;
                nr++;
                run_s = fs;
                run_n = (size_t)(fe - fs);
            }
            long long v;
            if (!field(p, e, start_col, &fs, &fe) || !parse_int(fs, fe, &v) || v < 0 || v > 2147483647LL)
                return fail(FPT_ERR_ARG, "GTrack record %lld: start is not a non-negative 32-bit integer", (long long)r);
            pos[r] = (int32_t)v;
            if (!field(p, e, value_col, &fs, &fe) || fs == fe || fe - fs > 63)
                return fail(FPT_ERR_ARG, "GTrack record %lld has no value", (long long)r);
            char tmp[64];
            memcpy(tmp, fs, (size_t)(fe - fs));
            tmp[fe - fs] = 0;
            char *stop;
            double d = strtod(tmp, &stop);
            if (stop == tmp || *stop) return fail(FPT_ERR_ARG, "GTrack record %lld: value '%s' is not a number", (long long)r, tmp);
            vals[r] = d;
            r++;
        }
        p = e0 < end ? e0 + 1 : end;
    }
    if (r != nrecords) return fail(FPT_ERR_ARG, "expected %lld GTrack records, parsed %lld", (long long)nrecords, (long long)r);
    *nruns = nr;
    return FPT_OK;
}

/* reference-layout float64 values -> compact int8 codes (3, 0, -3 kept; everything else, i.e. -10000, -> -128) */
extern "C" int fpt_compact_codes(const double *vals, int64_t n, int8_t *codes) {
    if ((!vals || !codes) && n > 0) return fail(FPT_ERR_ARG, "fpt_compact_codes: null argument");
    for (int64_t i = 0; i < n; i++) {
        double v = vals[i];
        codes[i] = v == 3.0 ? 3 : (v == -3.0 ? -3 : (v == 0.0 ? 0 : -128));
    }
    return FPT_OK;
}

#endif
