/*
 * bossgpu.h — C ABI of libbossgpu.so: the B200 (sm_100a) implementation of BOSS-RUNS' periodic
 * strategy update (coverage update -> site scoring -> benefit smoothing -> strategy derivation).
 *
 * Every entry point below replaces a piece of the reference's NumPy hot path; the citation after
 * "replaces:" is file:line in the upstream repository (goldman-gp-ebi/BOSS-RUNS @ a5b6af8).
 *
 * Conventions
 *   - plain C: pointers + sizes only; no C++/torch types cross this boundary.
 *   - every function returns 0 on success and a negative BOSSGPU_E* code on failure; the message
 *     is available from bossgpu_last_error() (thread-local). Nothing throws across the boundary.
 *   - the caller owns every host buffer; the library owns all device memory inside the handle.
 *   - a handle is bound to one CUDA device and one stream and is not thread-safe.
 *   - there is no CPU fallback: without a usable CUDA device bossgpu_create() fails.
 *
 * Vocabulary (the reference's): contig, site, barcode, bucket (20 000 sites), bin (100 sites),
 * window (2 000 sites, read-start distribution), strand 0 = forward, 1 = reverse.
 */
#ifndef BOSSGPU_H
#define BOSSGPU_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define BOSSGPU_ABI_VERSION 1

/* error codes */
#define BOSSGPU_OK            0
#define BOSSGPU_EINVAL       -1   /* bad argument (mirrors the reference's AssertionError / ValueError sites) */
#define BOSSGPU_ECUDA        -2   /* CUDA runtime failure (message holds cudaGetErrorString) */
#define BOSSGPU_ENOMEM       -3
#define BOSSGPU_EBASE        -4   /* a read base outside ACGT reached the scatter: IndexError upstream
                                     (boss/runs/reference.py:138-140 with codes from sequences.py:762-763) */
#define BOSSGPU_ESHAPE       -5   /* CIGAR does not span tend-tstart / qend-qstart: upstream AssertionError
                                     (sequences.py:732-733) or NumPy shape ValueError (sequences.py:785) */
#define BOSSGPU_ESTATE       -6   /* call sequence violated (e.g. update phases out of order) */
#define BOSSGPU_EEMPTY       -7   /* all benefits are zero: upstream `np.max` of an empty array raises
                                     ValueError (sequences.py:588) */
#define BOSSGPU_EPEER        -8   /* sharded update: a peer shard did not reach an exchange step in time */
#define BOSSGPU_ENOTC        -9   /* a bucket is on but update_params.tc is NaN: upstream has no `time_cost` before the first
                                     successful ReadlengthDist.update (readlengthdist.py:68) and raises AttributeError when
                                     update_wrapper reads it (core.py:192); switches are updated, no strategy is touched */

/* model constants of the reference (function defaults upstream, fixed here) */
#define BOSSGPU_BIN          100     /* downsampling window: reference.py:109,215 ; sequences.py:577 */
#define BOSSGPU_BUCKET       20000   /* reference.py:83 */
#define BOSSGPU_RSD_WINDOW   2000    /* readstartdist.py:13 */
#define BOSSGPU_FREEZE       30      /* sequences.py:419 */
#define BOSSGPU_N_PATTERNS   278256  /* C(34,5): count patterns with sum <= 29 */
#define BOSSGPU_N_STEPS      10      /* eta-1 pieces of the read-length staircase: readlengthdist.py:9,86 */
#define BOSSGPU_HIST_BINS    1088    /* |binary exponent| of benefit/max lies in [0,1075] */

typedef struct bossgpu_handle bossgpu_handle;

/* ---------------------------------------------------------------------------------------------
 * Construction. Replaces: Contig.__init__ state allocation (boss/runs/reference.py:71-118),
 * Scoring.__init__/init_score_array (boss/runs/sequences.py:335-393) and the per-contig wiring of
 * BossRuns.init (boss/runs/core.py:23-55).
 *
 * A handle holds one *shard*: an ordered list of segments, each a position range of one
 * non-rejected contig, in `contigs_filt` order. With one GPU every segment is a whole contig.
 * ------------------------------------------------------------------------------------------- */
typedef struct bossgpu_segment {
    int32_t contig;        /* index into the GLOBAL contigs_filt order */
    int32_t reserved;
    int64_t contig_len;    /* full length of that contig */
    int64_t start;         /* first site of this segment within the contig; multiple of BOSSGPU_BUCKET */
    int64_t len;           /* sites in this segment; start+len == contig_len or a multiple of BOSSGPU_BUCKET */
} bossgpu_segment;

typedef struct bossgpu_config {
    int32_t abi_version;        /* BOSSGPU_ABI_VERSION */
    int32_t device;             /* CUDA device ordinal */
    void*   stream;             /* cudaStream_t to launch on (NULL = the legacy default stream) */
    int32_t n_segments;
    int32_t n_barcodes;         /* 1 when the run is not barcoded (core.py:31-35) */
    const bossgpu_segment* segments;
    const uint8_t* ref_codes;   /* seq_int of each segment, concatenated (0..3; reference.py:46-68) */
    /* global geometry, identical on every shard */
    int32_t n_contigs_total;    /* len(contigs_filt) */
    int32_t halo_bins;          /* capacity of the bin halo kept on split contig edges (0 = none) */
    const int64_t* contig_len_all;  /* [n_contigs_total] */
    int64_t n_sites_total;      /* Reference.n_sites: includes 4 per reject ref (reference.py:337,343-347) */
    int64_t n_windows_total;    /* sum over contigs_filt of int(L/2000) (readstartdist.py:26-28) */
    /* scoring constants, computed by the host mirror of `Priors` so they are bit-identical */
    int32_t len_g;              /* 5 haploid / 15 diploid genotypes */
    int32_t reserved2;
    const double* phi;          /* [5][len_g]          sequences.py:39-155 */
    const double* priors;       /* [4][len_g]          sequences.py:186-313 */
    const double* phi_pow;      /* [5][len_g][30]      phi**k, sequences.py:159-168 */
    double score0_contig;       /* score of a never-observed site: the HAPLOID score0 whatever the ploidy
                                   (Reference._load_contigs builds Contig(ploidy=1): reference.py:319,334) */
    double entropy0_contig;
} bossgpu_config;

int  bossgpu_abi_version(void);
const char* bossgpu_last_error(void);
int  bossgpu_device_count(int* n);
int  bossgpu_create(const bossgpu_config* cfg, bossgpu_handle** out);
int  bossgpu_destroy(bossgpu_handle* h);
int  bossgpu_synchronize(bossgpu_handle* h);

/* ---------------------------------------------------------------------------------------------
 * Coverage update. Replaces: CoverageConverter.convert_records/_parse_cigar
 * (boss/runs/sequences.py:678-794) + Contig.increment_coverage (boss/runs/reference.py:122-144)
 * + BossRuns._effect_increments (boss/runs/core.py:77-86).
 *
 * ingest_packed: the batch is already tokenised. Read i maps to segment seg[i] starting at
 *   contig coordinate tstart[i] (= min(tstart,tend)); its CIGAR is cigar[cig_off[i]..cig_off[i+1])
 *   with each op packed as (len << 4) | class, class 0 = consumes read+reference (M,=,X and every
 *   other letter the reference does not special-case), 1 = insertion (read only), 2 = deletion
 *   (reference only; counted as base 4). bases[base_off[i]..base_off[i+1]) are the read bases of
 *   the aligned slice in ALIGNMENT orientation: ASCII if base_is_ascii, else codes 0..3.
 *   ASCII is translated like upstream: ACGT -> 0..3, anything else -> ord-48 (-> BOSSGPU_EBASE).
 *   Positions outside the segment are clipped (a read spanning a shard edge is given to both shards).
 *   All pointers are HOST pointers unless on_device != 0 (then they are device pointers on the
 *   handle's device and no copy is made).
 * ingest_records: text form. Host threads only copy the CIGAR text and pack the read bases 2 bits each into pinned
 *   staging (copies to the device overlap that work); the CIGARs are tokenised on the GPU (regex + translate
 *   upstream: sequences.py:672,762-776) and checked against tend-tstart and the slice length before any counter is
 *   touched (upstream's asserts, sequences.py:732-733,785 -> BOSSGPU_ESHAPE). seq slices are given in ORIGINAL read orientation together with
 *   rev[i]; the library reverse-complements (boss/utils.py:85-95: ATGC<->TACG only). `contig` is the GLOBAL
 *   index in contigs_filt order. Every shard is handed the WHOLE batch: the library keeps the reads that overlap
 *   one of its segments (a shard holds at most one segment per contig; the scatter clips at its edges, so a read
 *   spanning a shard edge is counted in both shards, each its own part) and advances every contig's depth total
 *   by the reference span of all the batch's reads on it, so the dropout rule sees contig-wide depth on every shard.
 * contig_cov_add[k] (may be NULL): number of reference positions this batch adds to GLOBAL contig k
 *   over all shards — needed by the dropout rule (reference.py:158,175-177) when a contig is split;
 *   NULL means "this shard sees every read of its contigs" and the library counts by itself.
 * ------------------------------------------------------------------------------------------- */
int bossgpu_ingest_packed(bossgpu_handle* h, int64_t n_reads,
                          const int32_t* seg, const int64_t* tstart, const int32_t* barcode,
                          const int64_t* cig_off, const uint32_t* cigar,
                          const int64_t* base_off, const uint8_t* bases,
                          int base_is_ascii, int on_device, const int64_t* contig_cov_add);

int bossgpu_ingest_records(bossgpu_handle* h, int64_t n_reads,
                           const int32_t* contig, const int64_t* tstart, const int64_t* tend,
                           const int32_t* barcode, const uint8_t* rev,
                           const int64_t* cig_off, const char* cigar_text,
                           const int64_t* seq_off, const char* seq_text,
                           int n_threads);

/* ingest_records_ptr: as ingest_records, but the text stays where the caller has it: cigar_ptr[i] points at
 * cigar_len[i] characters, seq_ptr[i] at the read's characters of which [seq_from[i], seq_to[i]) is the
 * aligned slice (original orientation). Lets a Python caller pass the buffers of its str objects without
 * joining or copying them. */
int bossgpu_ingest_records_ptr(bossgpu_handle* h, int64_t n_reads,
                               const int32_t* contig, const int64_t* tstart, const int64_t* tend,
                               const int32_t* barcode, const uint8_t* rev,
                               const uint64_t* cigar_ptr, const int64_t* cigar_len,
                               const uint64_t* seq_ptr, const int64_t* seq_from, const int64_t* seq_to,
                               int n_threads);

/* ingest_records_routed: one process per GPU, each holding a range of the genome. The caller has already routed the batch:
 * it hands over only the reads whose interval overlaps this shard's range of their contig (converting the others would be
 * wasted host work, N times over), plus batch_cov_add[k] = reference span of the WHOLE batch on global contig k — the
 * dropout rule (reference.py:157-158,175-177) needs contig-wide depth on every shard. An empty read list is fine. */
int bossgpu_ingest_records_routed(bossgpu_handle* h, int64_t n_reads,
                                  const int32_t* contig, const int64_t* tstart, const int64_t* tend,
                                  const int32_t* barcode, const uint8_t* rev,
                                  const uint64_t* cigar_ptr, const int64_t* cigar_len,
                                  const uint64_t* seq_ptr, const int64_t* seq_from, const int64_t* seq_to,
                                  const int64_t* batch_cov_add, int n_threads);

/* Multi-shard geometry and halo staging (see bossgpu_update_phase) */
/* Split score/bin pass (optional; results are identical with and without it).
 *   bossgpu_prescore_begin  when a batch arrives, before anything is known about it: every 2000-site tile is scored at
 *                           once on a second stream from the counters as they are, with the dropout thresholds of the
 *                           current depth totals (reference.py:157-158), while the host picks records, packs bases and
 *                           copies (bossgpu_ingest_records*).
 *   bossgpu_prescore        once the batch's alignment intervals are known (contig index, tstart, tend of every read's
 *                           winning record — what CoverageConverter.convert_records derives, sequences.py:694-731):
 *                           marks the tiles the batch will write to.
 * The next update then scores only the marked tiles, plus every tile of a contig whose threshold moved with the new
 * depth total; all other tiles keep the early pass' bins, depth totals and dropout counts, which are what the update
 * would compute (same counters, same threshold). Every tile is scored from its counters in every update — nothing is
 * carried over from one update to the next. If the batch that is ingested differs from the announced one (or is
 * rejected, or arrives by another ingest route, or was never announced) the update scores every tile again.
 * Works with up to 24 barcodes (one CTA takes a tile through every barcode: the row rules Q6/Q8 couple them); a no-op beyond. */
int bossgpu_prescore_begin(bossgpu_handle* h);
int bossgpu_prescore(bossgpu_handle* h, int64_t n_reads, const int32_t* contig, const int64_t* tstart, const int64_t* tend);

int bossgpu_set_shards(bossgpu_handle* h, int32_t n_shards, int32_t shard_index, const int64_t* row_start);
int bossgpu_halo_pack(bossgpu_handle* h);
int bossgpu_halo_unpack(bossgpu_handle* h);

/* Host tokenizer on its own (no device work): CIGAR text -> packed ops. Returns the number of ops
 * written (<= cap) or a negative error; ref_span/query_span receive the spans the ops consume. */
int64_t bossgpu_tokenize_cigar(const char* text, int64_t len, uint32_t* out, int64_t cap,
                               int64_t* ref_span, int64_t* query_span);

/* ---------------------------------------------------------------------------------------------
 * Strategy update. Replaces BossRuns.update_wrapper (boss/runs/core.py:160-198):
 *   Scoring.update_scores (sequences.py:398-455) + Contig.modify_scores (reference.py:148-179)
 *   + Contig.check_buckets (reference.py:183-211) + ReadStartDist._expand_fhat
 *   (readstartdist.py:121-152) + Contig.calc_smu/calc_u (reference.py:215-269)
 *   + Scoring.merge_benefit/adjust_length (sequences.py:553-560, utils.py:206-226)
 *   + Scoring.find_strat_thread (sequences.py:566-649) + BossRuns._distribute_strategy
 *   (core.py:125-155).
 * ------------------------------------------------------------------------------------------- */
typedef struct bossgpu_update_params {
    int32_t w[BOSSGPU_N_STEPS];       /* approx_ccl // 100 (reference.py:252); each must be >= 1 */
    double  mult[BOSSGPU_N_STEPS];    /* np.arange(0.05, 1, 0.1)[::-1] (reference.py:253) */
    double  tc;                       /* time_cost // 100 (sequences.py:581) */
    double  bucket_threshold;         /* optional.bucket_threshold (core.py:108) */
    const double* fhat_windows;       /* HOST [n_windows_total][2]: F-hat per 2 kb window before expansion
                                         (readstartdist.py:86-115); NULL keeps the previous upload */
    int32_t write_debug;              /* != 0: also keep S_mu and expected benefit for the getters */
    int32_t fhat_from_counts;         /* != 0: derive F-hat on the device from the counts added with
                                         bossgpu_read_starts_add, using the three scalars below
                                         (readstartdist.py:96-115): (alpha + C) / denom where C > 0, else zero_value */
    double  rs_alpha;
    double  rs_denom;                 /* 2 * n_windows * alpha + sum(C) */
    double  rs_zero_value;            /* (1 - p0' * B(alpha, ..+sum C) / B(alpha, ..)) * alpha / denom */
} bossgpu_update_params;

typedef struct bossgpu_update_result {
    int32_t switched_on;      /* any bucket of any contig on (core.py:110-111); 0 => strategy left as is */
    int32_t strat_size;       /* argmax+1 over the exponent bins (sequences.py:636) */
    double  threshold;        /* acceptance threshold (sequences.py:643-646) */
    double  normaliser;       /* max non-zero benefit (sequences.py:588) */
    double  ubar0;            /* sum(fhat * smu) with smu := benefit, as upstream (core.py:182-183) */
    double  fhat_sum;         /* sum of the expanded F-hat before normalisation (readstartdist.py:144) */
    int64_t n_nonzero;        /* benefit entries entering the histogram */
    int64_t n_dropout;        /* site rows zeroed by the dropout rule (reference.py:160) */
    int64_t n_accept[2];      /* accepted bins per strand over all contigs (core.py:152-153 log) */
    int64_t mirror_bytes;     /* bytes of the host strategy mirror rewritten by this update (only 512-byte pieces that changed move) */
} bossgpu_update_result;

/* single-shard update: all phases back to back on the handle's stream, one host sync at the end */
int bossgpu_update(bossgpu_handle* h, const bossgpu_update_params* p, bossgpu_update_result* r);

/* Multi-shard update = the same kernels split where the path has an exchange step. The caller
 * (one process per GPU) performs the exchanges on the exposed device buffers with NCCL:
 *   phase 0  score+bin pass, bucket switches         -> exchange: scores_ds halos of split contigs,
 *                                                        allreduce(max) of the switch flag
 *   phase 1  S_mu / staircase benefit, local max     -> allreduce(max) of the normaliser word
 *   phase 2  exponent histogram                      -> allreduce(sum) of the integer histogram
 *   phase 3  threshold, mask of own merged rows      -> allgather of the packed merged mask
 *   phase 4  bucket-gated distribution into the persistent strategy */
int bossgpu_update_phase(bossgpu_handle* h, int phase, const bossgpu_update_params* p,
                         bossgpu_update_result* r);

#define BOSSGPU_BUF_SWITCH     0   /* int32[1]                          allreduce max */
#define BOSSGPU_BUF_NORM       1   /* uint64[1] (double bits, >= 0)     allreduce max */
#define BOSSGPU_BUF_HIST       2   /* uint64[3*HIST_BINS + 4]           allreduce sum */
#define BOSSGPU_BUF_MASK       3   /* uint8[...] merged mask, all shards allgather    */
#define BOSSGPU_BUF_HALO_SEND  4
#define BOSSGPU_BUF_HALO_RECV  5
#define BOSSGPU_BUF_COV_TOTAL  7   /* uint64[n_contigs_total] depth total per contig: allreduce sum ONCE after
                                      bossgpu_set_coverage / bossgpu_synth_coverage on split contigs (ingest keeps it global) */
#define BOSSGPU_BUF_STRAT      6   /* uint8[strat rows * 2 * n_barcodes]: this shard's Contig.strat rows (gather to the
                                      process that writes boss.npz) */
int bossgpu_exchange_buffer(bossgpu_handle* h, int which, void** dev_ptr, size_t* bytes);

/* Peer-memory fabric: the same four exchanges done by the GPUs themselves, with no host round trip and no
 * library collective inside an update. Upstream has no counterpart (it is one process: the per-contig loops of
 * boss/runs/core.py:83-121 and the global threshold of sequences.py:566-649 run back to back); the exchange
 * points are the ones listed for bossgpu_update_phase.
 *
 * Every shard owns one exchange block in its HBM (allocated by bossgpu_set_shards). Peers map it — through the
 * CUDA IPC handle when shards are processes (one per GPU of the NVSwitch box), or simply by address when they
 * are handles of one process on one device ("virtual shards", each on its OWN stream) — and store their
 * contributions and an epoch flag into it over NVLink; a shard spins on its own block only.
 *   bossgpu_fabric_info    this shard's block: device pointer, size, 64-byte IPC handle (any out may be NULL)
 *   bossgpu_ipc_open/close map / unmap a peer process' block on `device`
 *   bossgpu_fabric_attach  peer_ptrs[n_shards]: every shard's block as THIS device addresses it (entry
 *                          shard_index = own block). timeout_s <= 0 keeps the default (2 s): a peer that does
 *                          not show up makes bossgpu_update_fused_end return BOSSGPU_EPEER instead of hanging.
 *   bossgpu_update_fused_begin  enqueue ALL kernels of one update, exchanges included; returns at once. Every
 *                          shard must call it the same number of times (the epoch is the call count).
 *   bossgpu_update_fused_end    one stream synchronisation, then the result record (as bossgpu_update). */
int bossgpu_fabric_info(bossgpu_handle* h, void** dev_ptr, size_t* bytes, unsigned char ipc_handle[64]);
int bossgpu_ipc_open(int device, const unsigned char ipc_handle[64], void** dev_ptr);
int bossgpu_ipc_close(int device, void* dev_ptr);
int bossgpu_fabric_attach(bossgpu_handle* h, int32_t n_shards, const uint64_t* peer_ptrs, double timeout_s);
int bossgpu_update_fused_begin(bossgpu_handle* h, const bossgpu_update_params* p);
int bossgpu_update_fused_end(bossgpu_handle* h, bossgpu_update_result* r);

/* ---------------------------------------------------------------------------------------------
 * Results and state access (host buffers, reference layouts).
 * ------------------------------------------------------------------------------------------- */
/* Contig.strat of the part of segment `seg` in this shard: bool [len//100 rows][2][n_barcodes]
 * (reference.py:118; read by simulation.py:79-80 and written to boss.npz by core.py:59-69). */
int bossgpu_get_strat(bossgpu_handle* h, int32_t seg, uint8_t* out, int64_t out_bytes);
/* every segment back to back, same layout, one copy */
int bossgpu_get_strat_all(bossgpu_handle* h, uint8_t* out, int64_t out_bytes);
/* same bits packed little-endian, 8 per byte, over the flattened [row][strand][barcode] order */
int bossgpu_get_strat_packed(bossgpu_handle* h, uint8_t* out, int64_t out_bytes);
int64_t bossgpu_strat_rows(bossgpu_handle* h, int32_t seg);   /* seg = -1: all segments */

/* Host mirror of every segment's strategy, back to back in the layout of bossgpu_get_strat_all (what
 * Contig.strat views, reference.py:118). The distribution kernel keeps it current by itself: it writes the 512-byte
 * pieces whose bytes changed straight into this (mapped, pinned) memory, so after bossgpu_update returns the
 * mirror equals the device state and callers wrap it once, zero-copy. Valid until bossgpu_destroy. */
int bossgpu_strat_host(bossgpu_handle* h, uint8_t** ptr, int64_t* bytes);
/* Use caller-provided host memory as that mirror instead (exactly bossgpu_strat_rows(h,-1)*2*n_barcodes bytes).
 * With one process per GPU every shard is given its slice of ONE shared-memory array, so the process that writes
 * boss.npz sees all masks without any gather (core.py:59-69). registered == 0: the library page-locks and maps
 * the pages around the range itself; != 0: the caller has done so for a range containing it (bossgpu_host_register,
 * once per process — slices of one array share pages). The current strategy is copied into it; the memory must
 * outlive the handle. */
int bossgpu_set_strat_mirror(bossgpu_handle* h, void* host_ptr, int64_t bytes, int registered);
/* cudaHostRegister (mapped, portable) / cudaHostUnregister of the pages around a host range, for callers without
 * a CUDA binding of their own. Page-locking works on whole pages: hand in memory that owns its pages (a shared-memory
 * segment, an anonymous mapping). A heap array shares its first and last page with unrelated allocations, and a later
 * pageable cudaMemcpy into such a neighbour (half inside a locked page) is refused by the driver. */
int bossgpu_host_register(void* host_ptr, int64_t bytes);
int bossgpu_host_unregister(void* host_ptr);
/* accepted entries per segment and strand after the last update, int64 [n_segments][2]
 * (numerators of the per-contig log line, core.py:152-154) */
int bossgpu_get_seg_accept(bossgpu_handle* h, int64_t* out, int64_t n);

/* Read-start counts on the device. Replaces the accumulation half of ReadStartDist.count_read_starts
 * (boss/runs/readstartdist.py:68-82): window = global index of the 2 kb window (contigs_filt order,
 * int(L/2000) windows per contig; entries outside [0, n_windows_total) are dropped like np.histogram drops
 * out-of-range starts), strand 0 forward / 1 reverse. */
int bossgpu_read_starts_add(bossgpu_handle* h, int64_t n, const int64_t* window, const uint8_t* strand);
int bossgpu_get_read_starts(bossgpu_handle* h, int64_t* out, int64_t n);   /* int64 [n_windows_total][2] */

/* Contig.coverage uint16 [len][5][n_barcodes] (reference.py:77) */
int bossgpu_get_coverage(bossgpu_handle* h, int32_t seg, uint16_t* out, int64_t out_elems);
int bossgpu_set_coverage(bossgpu_handle* h, int32_t seg, const uint16_t* in, int64_t in_elems);
/* Contig.scores / Contig.entropy float64 [len][n_barcodes] as they stand after the last update
 * (materialised on demand from the counts; entropy of frozen sites is the table value, see Q7) */
int bossgpu_get_scores(bossgpu_handle* h, int32_t seg, double* scores, double* entropy, int64_t out_elems);
/* Contig.scores_ds float64 [bins][n_barcodes] (reference.py:227) */
int bossgpu_get_scores_ds(bossgpu_handle* h, int32_t seg, double* out, int64_t out_elems);
/* Contig.additional_benefit / smu / expected_benefit float64 [bins][2][n_barcodes]
 * (reference.py:225,254,267); smu and expected need write_debug in the last update */
int bossgpu_get_benefit(bossgpu_handle* h, int32_t seg, double* additional, double* smu,
                        double* expected, int64_t out_elems);
/* Contig.bucket_switches bool [len//20000+1][n_barcodes] and switched_on bool [n_barcodes] */
int bossgpu_get_buckets(bossgpu_handle* h, int32_t seg, uint8_t* switches, int64_t n, uint8_t* switched_on);
int bossgpu_set_buckets(bossgpu_handle* h, int32_t seg, const uint8_t* switches, int64_t n);
/* pinned host image of every segment's switches back to back ([n_sw][n_barcodes] each), refreshed at the end of
 * every update: wrap once, zero-copy */
int bossgpu_buckets_host(bossgpu_handle* h, uint8_t** ptr, int64_t* bytes);
/* exponent histogram of the last update: counts int64[HIST_BINS], f_grid float64[HIST_BINS]
 * (sequences.py:593-624, before the empty bins are dropped) */
int bossgpu_get_hist(bossgpu_handle* h, int64_t* counts, double* f_grid);
/* F-hat as the last update's histogram consumed it: float64 [n_rows][2] for rows [row0, row0 + n_rows) of the merged,
 * length-adjusted axis (expansion x20 + both tail fixes + normalisation, readstartdist.py:121-152, core.py:184-185;
 * identical for every barcode, core.py:175). Rows >= n_sites_total // 100 read 0. */
int bossgpu_get_fhat(bossgpu_handle* h, int64_t row0, int64_t n_rows, double* out);
/* the dense score table: float64 [N_PATTERNS][4] (replaces score_arr / entropy_arr, sequences.py:387-388) */
int bossgpu_get_score_table(bossgpu_handle* h, double* scores, double* entropies);
/* rank of a count pattern (c0..c4, sum <= 29) in that table; -1 if out of range */
int64_t bossgpu_pattern_rank(const uint16_t c[5]);

/* per-kernel device times of the last update in ms (CUDA events on the handle's stream):
 * [0] scatter (last ingest) [1] score+bin pass [2] buckets [3] smoothing [4] histogram
 * [5] threshold [6] mask+distribute [7] whole update */
#define BOSSGPU_N_TIMERS 8
int bossgpu_timing(bossgpu_handle* h, float ms[BOSSGPU_N_TIMERS]);
/* algorithmic launches issued so far (kernels of this library only) */
int64_t bossgpu_launch_count(bossgpu_handle* h);
/* host->device bytes of the last text ingest (per-read scalars + 4-byte CIGAR ops + read bases packed 2 bits each) */
int64_t bossgpu_ingest_bytes(bossgpu_handle* h);

/* Bench/test support: fill the coverage of every segment with a synthetic sequencing state on the
 * device (depth ~ Poisson(mean_depth) split over bases with p_ref, uniform errors, p_del; a fraction
 * of 20 kb regions with zero depth and a fraction with depth >= 30), deterministic in `seed`.
 * Not part of the reference; used for BASELINE.json config 3 whose state cannot be built on a host. */
int bossgpu_synth_coverage(bossgpu_handle* h, uint64_t seed, double mean_depth, double p_ref,
                           double p_del, double frac_dropout, double frac_deep);

/* ---------------------------------------------------------------------------------------------
 * BOSS-AEONS (no reference genome: the "contigs" are the current assembly and change with every batch, so this entry
 * point is stateless — no handle). Replaces, for a pool of contigs given as per-node scores (one per 100-bp node):
 *   Benefit.calc_fragment_benefit + helpers (boss/aeons/sequences.py:1555-1641), Benefit.benefit_bins (:1644-1682),
 *   ContigPool.find_threshold (:1059-1094) and Sequence.find_strat_m0 (:398-406).
 * node_off[n_seq + 1]: offsets of the contigs in `scores`; e1 / e2: left / right end markers (Sequence.noi[0], noi[-1]).
 * Outputs (host, caller-owned): benefit — contig i's forward row at 2 * node_off[i], its reverse row right behind it
 * (upstream's (2, n_i) array, flattened); smu_sum[n_seq]; strat — bool [total nodes][2]; counts[BOSSGPU_HIST_BINS] — the
 * exponent histogram. With want_strategy == 0 only benefit and smu_sum are produced (calc_fragment_benefit alone).
 * Errors: BOSSGPU_EEMPTY when every benefit is zero (np.max of an empty array upstream), BOSSGPU_EINVAL for windows < 1
 * (bn.move_sum) or decreasing windows.
 * ------------------------------------------------------------------------------------------- */
typedef struct bossgpu_aeons_params {
    int32_t mu_ds;                    /* mu // node_size */
    int32_t ccl_ds[BOSSGPU_N_STEPS];  /* approx_ccl // node_size */
    double  perc[BOSSGPU_N_STEPS];    /* np.arange(0.1, 1.1, 0.1)[::-1] (sequences.py:1634) */
    double  tc;                       /* (lam - mu - 300) // node_size (sequences.py:1074) */
    double  tbar0;                    /* alpha + rho + mu // node_size = 2 + 3 + 4 (sequences.py:1072-1073,1079) */
    int32_t want_strategy;
    int32_t reserved;
} bossgpu_aeons_params;

typedef struct bossgpu_aeons_result {
    double  threshold;
    double  normaliser;
    double  ubar0;                    /* sum of the contigs' smu_sum */
    int64_t n_nonzero;
    int32_t strat_size;
    int32_t reserved;
} bossgpu_aeons_result;

int bossgpu_aeons_update(int device, int64_t n_seq, const int64_t* node_off, const double* scores, const uint8_t* e1,
                         const uint8_t* e2, const bossgpu_aeons_params* p, double* benefit, double* smu_sum, uint8_t* strat,
                         int64_t* counts, bossgpu_aeons_result* r);

#ifdef __cplusplus
}
#endif
#endif /* BOSSGPU_H */
