/*
 * crgpu.h -- C ABI of libcrgpu.so: the B200 (sm_100a) implementation of CRISPResso's
 * read->amplicon alignment + indel quantification hot path.
 *
 * The reference (tonyreina/CRISPResso) has no FFI; the path is three in-process seams in
 * CRISPResso/CRISPRessoCORE.py::run_crispresso plus one subprocess:
 *
 *   S1  quality filter      filter_se_fastq_by_qual / filter_pe_fastq_by_qual /
 *                           get_ids_reads_to_remove            CORE:162-310, call site 1547-1583
 *   S2  alignment           `needle` subprocess + parse_needle_output
 *                                                               CORE:1791-1828, 1707-1786, 1911-1936
 *   S3  quantification      process_df_chunk                    CORE:428-753 (driver 2773-2864)
 *
 * Every entry point below names the seam it replaces.  Conventions:
 *   - plain C, no torch / C++ types; `int` return: 0 = ok, non-zero = CRGPU_E_* (message via
 *     crgpu_last_error).  The Python shim maps CRGPU_E_ALIGN to NeedleException (CLI exit 6,
 *     CORE:4349-4356).
 *   - the caller owns every buffer.  `mem` says where the caller's buffers live:
 *     CRGPU_MEM_HOST (library stages them through pinned memory, copies are part of the call)
 *     or CRGPU_MEM_DEVICE (device pointers on the context's GPU; nothing is copied).
 *   - one opaque context per GPU; a context is not thread-safe.
 *   - there is NO CPU fallback: without a CUDA device crgpu_create fails.
 */
#ifndef CRGPU_H
#define CRGPU_H

#include <stdint.h>
#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

#define CRGPU_ABI_VERSION 5

enum {
    CRGPU_OK = 0,
    CRGPU_E_CUDA = 1,       /* CUDA runtime error (no device, OOM, launch failure) */
    CRGPU_E_ARG = 2,        /* bad argument (null pointer, negative size, bad enum) */
    CRGPU_E_ALIGN = 3,      /* input the aligner cannot honour exactly: a base outside ACGTN(U),
                               read/amplicon length out of range, gap penalties not exactly
                               representable -> NeedleException in the shim */
    CRGPU_E_NOMEM = 4
};

enum { CRGPU_MEM_HOST = 0, CRGPU_MEM_DEVICE = 1 };

/* Limits of the exact int16x2 aligner (DESIGN.md "Score range"). */
#define CRGPU_MAX_AMPLICON 1024
#define CRGPU_MAX_READ 2048
#define CRGPU_MIN_LEN 2

typedef struct crgpu_ctx crgpu_ctx;

/* ---- context -------------------------------------------------------------------------- */
int crgpu_abi_version(void);
int crgpu_create(crgpu_ctx **ctx, int device);
void crgpu_destroy(crgpu_ctx *ctx);
const char *crgpu_last_error(const crgpu_ctx *ctx);
/* Cap on the traceback scratch held in HBM per batch (bytes; default: 24 GiB or an eighth of the device's memory,
 * whichever is less; there are two such sets).  A call runs as two batches -- the walks of one beside the fills of the
 * other -- or as many as this cap needs; results do not depend on it. */
int crgpu_set_traceback_budget(crgpu_ctx *ctx, size_t bytes);
/* Traceback walks of batch b normally run on a second stream, overlapped with the fill of batch b+1
 * (default on).  Turning it off serialises the kernels, which is what per-kernel timing wants. */
int crgpu_set_overlap(crgpu_ctx *ctx, int on);
/* The HDR-amplicon pass normally reuses the DP rows it shares with the amplicon pass (bit-identical
 * results, fewer cells evaluated; default on).  Off = two full passes. */
int crgpu_set_share_prefix(crgpu_ctx *ctx, int on);
/* Banded two-pass fill (default half-width 16 read columns; 0 = single-pass fill with flags for every
 * cell).  The first pass evaluates every DP cell without traceback flags, the second re-evaluates, with
 * flags, only the columns within half_width of the main diagonal of each lane's rows; a read whose
 * traceback leaves the band is re-aligned with the single-pass fill, so results never depend on this.  When
 * more than a quarter of a call's reads leave the band, the next 8 fused calls on the context use the
 * single-pass fill. */
int crgpu_set_band(crgpu_ctx *ctx, int half_width);
/* The current half-width (-1: no context). */
int crgpu_get_band(const crgpu_ctx *ctx);
/* Reads the last crgpu_align_quantify call re-aligned with the single-pass fill because their traceback
 * left the band: out[0] amplicon pass, out[1] HDR-amplicon pass. */
int crgpu_last_escaped(const crgpu_ctx *ctx, int out[2]);
/* Diagonal shortcut of the banded fill (default on).  Right after the score pass an alignment whose start-cell score
 * equals the sum of the substitution scores along the diagonal through the start cell is emitted directly: needle's
 * traceback is then provably that diagonal (DESIGN.md "Diagonal shortcut"), so only the remaining read pairs go through
 * the band pass and the walk.  Results never depend on this. */
int crgpu_set_diag_shortcut(crgpu_ctx *ctx, int on);
/* Exact-read shortcut of crgpu_align_quantify (default on): a read that IS the amplicon (same length and bases, case aside;
 * amplicon of A C G T only) aligns to it along the diagonal with the maximum possible score, identity 100.0, and to the HDR
 * amplicon like every other such read -- so all but one representative skip the DP; their records, ops and text rows are
 * written directly.  Results do not depend on it.  crgpu_last_exact: reads that skipped the DP in the last call. */
int crgpu_set_exact_shortcut(crgpu_ctx *ctx, int on);
int64_t crgpu_last_exact(const crgpu_ctx *ctx);
/* Last crgpu_align_quantify call: out[0] = read pairs of the banded passes, out[1] = pairs that still needed the band
 * pass + walk (equal when the shortcut is off). */
int crgpu_last_diag(const crgpu_ctx *ctx, int64_t out[2]);
/* Device time (ms, CUDA events on the context's stream) spent in each kernel family during
 * the LAST call on this context, and launch counts.  out_ms[0..5] = encode, fill, walk,
 * quantify, qualfilter, other;  out_launches likewise. */
int crgpu_last_timing(const crgpu_ctx *ctx, float out_ms[6], int64_t out_launches[6]);
/* The fill slot of crgpu_last_timing split by kernel kind: [0] single-pass fill (flags for every cell),
 * [1] score pass and [2] band pass of the banded fill -- device time (ms), launches and the DP cells
 * each kind evaluated during the LAST call (the numerators of the per-kernel rooflines). */
int crgpu_last_fill_breakdown(const crgpu_ctx *ctx, double out_ms[3], int64_t out_launches[3], int64_t out_cells[3]);
/* Synchronise the context's stream. */
int crgpu_sync(crgpu_ctx *ctx);
/* The context's cudaStream_t (as void*), so that a caller can bracket calls with its own CUDA
 * events on the stream the kernels are launched on. */
void *crgpu_stream(crgpu_ctx *ctx);

/* ---- S1: quality filter (CORE:162-193, 270-310) --------------------------------------- *
 * keep[i] = 1 iff mean(phred) >= min_mean_q and min(phred) >= min_single_q, phred = byte - 33,
 * evaluated in exact integer form (sum >= q*len).  For paired ends the caller ANDs the two
 * masks (CORE:216-227 drops a pair if either mate fails). An empty read has keep = 0. */
int crgpu_qualfilter(crgpu_ctx *ctx, int mem, const uint8_t *qual, const int64_t *offsets, int64_t n,
                     int min_mean_q, int min_single_q, uint8_t *keep);

/* ---- S2: alignment (needle subprocess + parse_needle_output, CORE:1791-1806, 1707-1786) -- */
typedef struct {
    float score;          /* needle "# Score:" (exact: a multiple of 1/8 far below 2^20) */
    int32_t alnlen;       /* alignment columns incl. end gaps ("# Length:") */
    int32_t ident;        /* identical columns ("# Identity: ident/alnlen") */
    int32_t tenths;       /* identity %% as printed by "%4.1f", times 10: what the parser's
                             score_<name> column holds (CORE:1732-1738) */
    int32_t aln_off;      /* the three strings occupy [aln_off, aln_off+alnlen) of the read's slot */
    int32_t start1;       /* traceback start cell: amplicon index */
    int32_t start2;       /* traceback start cell: read index */
    int32_t read_len;     /* the parser's `length` column (CORE:1754) */
} crgpu_aln_rec;

/* Align each of n reads (bytes reads[offsets[i]..offsets[i+1])) globally to `amplicon` exactly
 * as `needle -gapopen G -gapextend E` (EDNAFULL, end gaps free) does, including its traceback
 * tie-breaking.  recs[n] always written.  ref_out/mark_out/qry_out (each n*slot bytes, slot >=
 * amplicon_len + longest read) receive the three srspair rows -- aligned amplicon, markup
 * ('|' '.' ' '), aligned read -- right-aligned in the read's slot (see aln_off); pass NULL for
 * all three to skip them (score-only use, CORE:1833-1835 just_score=True).
 * Reads shorter than CRGPU_MIN_LEN or with a base outside ACGTN(U) fail the call (CRGPU_E_ALIGN). */
int crgpu_align(crgpu_ctx *ctx, int mem, const char *amplicon, int amplicon_len,
                const uint8_t *reads, const int64_t *offsets, int64_t n,
                double gapopen, double gapextend,
                crgpu_aln_rec *recs, uint8_t *ref_out, uint8_t *mark_out, uint8_t *qry_out, int64_t slot);

/* ---- S3: quantification (process_df_chunk, CORE:428-753) ------------------------------- */
enum {                        /* crgpu_quant_params.flags */
    CRGPU_Q_HAS_HDR = 1,          /* args.expected_hdr_amplicon_seq non-empty (CORE:537) */
    CRGPU_Q_IGNORE_SUBS = 2,      /* CORE:490 */
    CRGPU_Q_IGNORE_INS = 4,       /* CORE:517 */
    CRGPU_Q_IGNORE_DEL = 8,       /* CORE:503 */
    CRGPU_Q_WINDOW = 16,          /* args.window_around_sgrna != 0 (CORE:611) */
    CRGPU_Q_HIDE_OUTSIDE = 32,    /* args.hide_mutations_outside_window_NHEJ (CORE:591,645) */
    CRGPU_Q_FRAMESHIFT = 64,      /* args.coding_seq non-empty (CORE:446) */
    CRGPU_Q_MASK_N = 128          /* amplicon contains N: apply ignore_n_in_alignment (CORE:2033-2052) */
};

enum {                        /* class bits in crgpu_read_rec.cls (DataFrame columns CORE:2014-2019) */
    CRGPU_C_UNMODIFIED = 1, CRGPU_C_NHEJ = 2, CRGPU_C_HDR = 4, CRGPU_C_MIXED = 8
};

/* order of the 15 per-position vectors in `vectors` (process_df_chunk's return tuple, CORE:730-753) */
enum {
    CRGPU_V_INS = 0, CRGPU_V_DEL, CRGPU_V_MUT, CRGPU_V_ANY,
    CRGPU_V_INS_MIXED, CRGPU_V_DEL_MIXED, CRGPU_V_MUT_MIXED,
    CRGPU_V_INS_HDR, CRGPU_V_DEL_HDR, CRGPU_V_MUT_HDR,
    CRGPU_V_INS_NONCODING, CRGPU_V_DEL_NONCODING, CRGPU_V_MUT_NONCODING,
    CRGPU_V_AVG_DEL, CRGPU_V_AVG_INS,
    CRGPU_NUM_VECTORS
};
enum { CRGPU_K_MOD_FRAMESHIFT = 0, CRGPU_K_MOD_NON_FRAMESHIFT, CRGPU_K_NON_MOD_NON_FRAMESHIFT,
       CRGPU_K_SPLICING_MODIFIED, CRGPU_NUM_COUNTERS };

typedef struct {
    int32_t amplicon_len;                  /* LEN_AMPLICON (CORE:1288) */
    int32_t flags;                         /* CRGPU_Q_* */
    double hdr_perfect_alignment_threshold;/* CORE:541 */
    const uint8_t *include_mask;           /* [amplicon_len] 1 = in INCLUDE_IDXS (CORE:2739-2762); host memory */
    const uint8_t *exon_mask;              /* [amplicon_len] EXON_POSITIONS (CORE:1418-1451) or NULL; host */
    const uint8_t *splice_mask;            /* [amplicon_len] SPLICING_POSITIONS (CORE:1444-1455) or NULL; host */
} crgpu_quant_params;

typedef struct {
    uint8_t cls;          /* CRGPU_C_* after classification */
    uint8_t pad[3];
    int32_t n_mutated;    /* CORE:654 */
    int32_t n_inserted;   /* CORE:657 */
    int32_t n_deleted;    /* CORE:660 */
} crgpu_read_rec;

/* Quantify n aligned reads.  Strings as produced by crgpu_align: row i occupies
 * X[i*slot + aln_off[i] .. +alnlen[i]) (aln_off may be all zeros for left-aligned rows).
 * tenths_ref[i] / tenths_rep[i]: identity tenths vs the amplicon / the HDR amplicon
 * (tenths_rep < 0 = NaN: the reverse-complement rows of CORE:1924-1949).  unmodified_in[i] =
 * the UNMODIFIED column on entry (CORE:2014, 2047).
 * Outputs: out_recs[n]; vectors[CRGPU_NUM_VECTORS][amplicon_len] int64 ADDED to the caller's
 * values; hist_inframe / hist_frameshift: int64[hist_len] ADDED, bin index = key + hist_zero
 * (key = effective exon length change, CORE:710-717); counters[CRGPU_NUM_COUNTERS] ADDED.
 * vectors/hist/counters are always HOST memory (small); per-read arrays follow `mem`. */
int crgpu_quantify(crgpu_ctx *ctx, int mem, const crgpu_quant_params *params,
                   const uint8_t *ref_rows, const uint8_t *mark_rows, const uint8_t *qry_rows, int64_t slot,
                   const int32_t *aln_off, const int32_t *alnlen,
                   const int32_t *tenths_ref, const int32_t *tenths_rep, const uint8_t *unmodified_in,
                   int64_t n, crgpu_read_rec *out_recs,
                   int64_t *vectors, int64_t *hist_inframe, int64_t *hist_frameshift, int32_t hist_len,
                   int32_t hist_zero, int64_t *counters);

/* ---- fused hot path: S2 (+HDR pass, + reverse-complement rescue) + S3 in one call -------- *
 * Reproduces CORE:1791-2072 + 2773-2864 for one amplicon without materialising the needle text:
 *  1. align all reads to `amplicon` (and to `hdr_amplicon` when non-NULL, CORE:1810-1828);
 *  2. reads with score_ref < min_identity are re-aligned to the reverse complement of the
 *     amplicon (CORE:1865-1867, 1873-1921; in HDR mode their score_repaired is NaN because the
 *     reference's repair-RC needle command is malformed, CORE:1924-1949 / SURVEY Q11);
 *  3. forward rows are kept iff score_ref > min_identity (HDR mode: or score_repaired >
 *     min_identity, CORE:1849-1852, 1869-1871); RC rows iff score_ref(rc) > min_identity
 *     (CORE:1956-1959, 1976-1978);
 *  4. UNMODIFIED = (score_ref == 100), N-masking when the amplicon has N (CORE:2014-2052), then
 *     every kept row is quantified with crgpu_quantify semantics.
 * A read can own a forward row AND an RC row (HDR mode only). */
typedef struct {
    double gapopen, gapextend;             /* parsed from --needle_options_string (CORE:4226-4231) */
    double min_identity_score;             /* CORE:4092 */
    const char *hdr_amplicon;              /* expected HDR amplicon or NULL */
    int32_t hdr_amplicon_len;
    int32_t rc_rescue;                     /* 1 = run the reverse-complement rescue (reference behaviour) */
} crgpu_path_params;

typedef struct {
    /* forward rows, indexed by read; per-read arrays follow `mem` */
    uint8_t *kept;                         /* [n] bit0: forward row kept, bit1: RC row kept,
                                              bit2: read was re-aligned to the reverse complement,
                                              bit3: not aligned -- a base outside ACGTN(U) (alone: no other bit is set) */
    crgpu_aln_rec *aln;                    /* [n] alignment vs the amplicon */
    int32_t *tenths_rep;                   /* [n] identity tenths vs the HDR amplicon (-1 = NaN / no HDR); may be NULL */
    crgpu_read_rec *recs;                  /* [n] valid where kept&1 */
    uint8_t *ref_rows, *mark_rows, *qry_rows;   /* n*slot each (right-aligned, see aln_off) or all NULL */
    int64_t slot;                          /* >= amplicon_len + longest read; 0 = choose (only without rows) */
    /* reverse-complement rescue rows: compact, in read order */
    int64_t rc_cap;                        /* capacity (rows) of the rc_* arrays */
    int64_t rc_n;                          /* OUT: number of reads re-aligned to the reverse complement */
    int32_t *rc_read;                      /* [rc_cap] read index of RC row j */
    crgpu_aln_rec *rc_aln;                 /* [rc_cap] alignment vs revcomp(amplicon) */
    crgpu_read_rec *rc_recs;               /* [rc_cap] valid where kept[rc_read[j]]&2 */
    uint8_t *rc_ref_rows, *rc_mark_rows, *rc_qry_rows;  /* rc_cap*slot each or NULL: rows already flipped to
                                              the forward strand (CORE:1982-1990), LEFT-aligned in the slot */
    /* reductions: HOST memory, results are ADDED to the caller's values */
    int64_t *vectors;                      /* [CRGPU_NUM_VECTORS][amplicon_len] */
    int64_t *hist_inframe, *hist_frameshift;    /* [hist_len], bin = key + hist_zero; may be NULL without -c */
    int32_t hist_len, hist_zero;
    int64_t *counters;                     /* [CRGPU_NUM_COUNTERS] */
    int64_t class_counts[4];               /* OUT (added): UNMODIFIED, NHEJ, HDR, MIXED rows (CORE:2866-2869) */
    int64_t n_total;                       /* OUT (added): rows kept = df_needle_alignment.shape[0] (CORE:2025) */
    int64_t n_cells;                       /* OUT (added): DP cells computed, for GCUPS */
    /* allele table (CORE:2923-2946): kept rows grouped by (align_seq, ref_seq, class, n_deleted,
     * n_inserted, n_mutated) on the device; HOST arrays, most frequent allele first */
    int64_t allele_cap;                    /* capacity of the two arrays below; 0 = skip the grouping */
    int64_t allele_n;                      /* OUT: distinct alleles among the kept rows (may exceed allele_cap) */
    int32_t *allele_row;                   /* [allele_cap] a representative row: read index i for a forward row,
                                              n + j for the RC row j */
    int64_t *allele_count;                 /* [allele_cap] #Reads */
    int64_t n_cells_computed;              /* OUT (added): DP cells actually evaluated.  Less than n_cells when the HDR
                                              pass reuses the DP rows it shares with the amplicon pass (same-length HDR
                                              amplicon: all rows above the first differing base are identical) */
    uint64_t *allele_key;                  /* [2 * allele_cap] or NULL; HOST: two independent 64-bit hashes of allele k's
                                              grouping key.  A caller that quantifies one read set in several calls (chunks)
                                              merges the per-call tables by them (hotpath.run_hot_path_pipelined) */
} crgpu_path_out;

int crgpu_align_quantify(crgpu_ctx *ctx, int mem, const char *amplicon, int amplicon_len,
                         const crgpu_path_params *path, const crgpu_quant_params *quant,
                         const uint8_t *reads, const int64_t *offsets, int64_t n, crgpu_path_out *out);

/* ---- staged inputs: the NEXT batch of reads is copied while the current one is computed ------- *
 * crgpu_stage_reads registers a batch for the host-to-device copy into staging slot 0 or 1 (on the context's copy
 * stream) and returns at once; `reads` / `offsets` must stay valid and unchanged -- and should be pinned -- until the
 * batch has been consumed.  The copy itself starts inside the next crgpu_align_quantify* call of the context, behind its
 * first kernel launches (or when the batch is run / crgpu_sync is called before that); crgpu_align_quantify_staged waits
 * for it and runs crgpu_align_quantify on the batch, with every output in HOST memory.  Typical loop: stage(0, b0); for k: stage((k+1)&1, b[k+1]); align_quantify_staged(k&1).
 * format: CRGPU_READS_BYTES = one base per byte, as everywhere else; CRGPU_READS_BAM4 = two bases per byte in BAM's
 * 4-bit codes "=ACMGRSVTWYHKDBN", high nibble first, dense (base j of the batch is nibble j; offsets count BASES):
 * half the bytes over PCIe, unpacked on the device.  Codes other than A C G T N are reported per read (kept bit 3). */
enum { CRGPU_READS_BYTES = 0, CRGPU_READS_BAM4 = 1 };
/* on = 1: crgpu_align_quantify_staged returns as soon as the reductions, the RC-rescue rows and the allele table are in
 * host memory; the per-read arrays (kept, aln, recs, tenths_rep) follow on the copy stream, behind the next batch's
 * kernels, and are valid after crgpu_sync (or after the staged call two batches later has returned). */
int crgpu_set_deferred_outputs(crgpu_ctx *ctx, int on);
int crgpu_stage_reads(crgpu_ctx *ctx, int slot, int format, const uint8_t *reads, const int64_t *offsets, int64_t n);
int crgpu_align_quantify_staged(crgpu_ctx *ctx, int slot, const char *amplicon, int amplicon_len,
                                const crgpu_path_params *path, const crgpu_quant_params *quant, crgpu_path_out *out);

/* ---- S0: paired-end merge -- the `flash` subprocess of CORE:1655-1664 -------------- *
 * `flash R1 R2 --allow-outies --max-overlap M --min-overlap m` (FLASH 1.2.11; default maximum mismatch
 * density 0.25).  Mate i of file 1 is bytes seq1/qual1[off1[i]..off1[i+1]) (phred+33 qualities share the
 * sequence offsets), likewise file 2.  Read 2 is reverse-complemented; the best overlap is the first
 * lexicographic minimum of (mismatch density, mismatch quality score) in FLASH's scan order (innies by
 * increasing start in read 1, then -- with allow_outies -- outies); a pair whose best density exceeds
 * max_mismatch_density stays uncombined.  Bases must be ACGTN, mates at most 1024 long (CRGPU_E_ALIGN). */
typedef struct {
    int32_t min_overlap;           /* args.min_paired_end_reads_overlap (CORE:4147), flash -m */
    int32_t max_overlap;           /* args.max_paired_end_reads_overlap (CORE:4140), flash -M */
    float max_mismatch_density;    /* flash -x, default 0.25 (CRISPResso never changes it) */
    int32_t allow_outies;          /* flash -O (CORE:1659 always passes it) */
} crgpu_merge_params;

typedef struct {
    /* per pair; follow `mem` */
    int32_t *pos;                  /* [n] start of the overlap in the left read (read 1 for an innie,
                                      revcomp(read 2) for an outie); -1 = not combined */
    uint8_t *kind;                 /* [n] 0 = not combined (out.notCombined_*), 1 = innie, 2 = outie */
    /* merged reads (out.extendedFrags), compact and in pair order; follow `mem` */
    int64_t cap_bytes, cap_reads;  /* capacities of seq/qual and of offsets(-1)/index; sum of all mate
                                      lengths and n always suffice */
    uint8_t *seq, *qual;           /* [cap_bytes] merged bases / phred+33 qualities */
    int64_t *offsets;              /* [cap_reads + 1] merged read j is bytes offsets[j]..offsets[j+1]: the
                                      (reads, offsets) pair crgpu_align / crgpu_align_quantify consume */
    int32_t *index;                /* [cap_reads] pair index of merged read j */
    int64_t n_merged;              /* OUT */
    int64_t n_innie, n_outie;      /* OUT: merged pairs by kind (FLASH's "innie" / "outie" counts) */
    int64_t bytes;                 /* OUT: offsets[n_merged] */
} crgpu_merge_out;

int crgpu_flash_merge(crgpu_ctx *ctx, int mem, const uint8_t *seq1, const uint8_t *qual1, const int64_t *off1,
                      const uint8_t *seq2, const uint8_t *qual2, const int64_t *off2, int64_t n,
                      const crgpu_merge_params *params, crgpu_merge_out *out);

/* ---- FASTQ record splitting ------------------------------------------------------ *
 * Indexes inflated FASTQ text: what `gunzip | awk 'NR%4==1 {...} NR%4==2 {...}'` (CORE:1793-1797), `wc -l`
 * (get_n_reads_fastq, CORE:335-348), the read-length awk (get_average_read_length_fastq, CORE:313-332)
 * and the Biopython record iterator of the quality filter (CORE:176-190) do on the host today.
 * text[0..nbytes) holds 4-line records ('\n' or '\r\n' line ends).  With final_chunk = 0 the text may
 * end inside a record: n_records complete records are indexed and `consumed` says where the incomplete
 * tail starts (prepend it to the next chunk).  With final_chunk = 1 a missing last newline is accepted and
 * a line count that is not a multiple of 4 is an error.  A record whose line 1 does not start with '@',
 * line 3 with '+', or whose bases and qualities differ in length fails the call (CRGPU_E_ARG).
 * Pass seq = qual = offsets = NULL to get the three counts only (then size the buffers exactly). */
typedef struct {
    int64_t n_records;             /* OUT: complete records = get_n_reads_fastq */
    int64_t consumed;              /* OUT: bytes of text those records occupy */
    int64_t seq_bytes;             /* OUT: sum of read lengths; seq_bytes / n_records (integer) =
                                      get_average_read_length_fastq */
    int64_t cap_records, cap_bytes;/* capacities of the arrays below */
    uint8_t *seq, *qual;           /* [cap_bytes] packed bases / phred+33 qualities; follow `mem` */
    int64_t *offsets;              /* [cap_records + 1]; follow `mem` */
    int64_t *name_start;           /* [cap_records] offset of the record's '@' in text, or NULL; follow `mem` */
    int32_t *name_len;             /* [cap_records] length of the header line incl. '@', or NULL */
} crgpu_fastq_out;

int crgpu_fastq_index(crgpu_ctx *ctx, int mem, const uint8_t *text, int64_t nbytes, int final_chunk,
                      crgpu_fastq_out *out);

/* ---- measurement helper ---------------------------------------------------------------- *
 * Integer issue-rate micro-benchmark ("measure it"): dependency-free chains of ONE
 * instruction kind on every SM.  Returns lane-ops per second (one SASS instruction on one lane).
 * which: 0 = IADD (alu pipe), 1 = IMAD (fma pipe), 2 = VIMNMX.S16x2, 3 = VIADDMNMX.S16x2,
 * 4 = 1:1 mix of VIMNMX.S16x2 (alu) and IMAD (fma) -- the dual-pipe integer peak the roofline
 * uses --, 5 = VIMNMX3.S16x2, 6 = LOP3, 7 = (sub, VIMNMX imm, add) triple. */
int crgpu_int_peak(crgpu_ctx *ctx, int which, double *lane_ops_per_s);

#ifdef __cplusplus
}
#endif
#endif /* CRGPU_H */
