/* minicom_b200.h — C-ABI of the B200-native minicom front end.
 *
 * One shared library, libminicom_b200.so (hand-written sm_100a CUDA behind
 * `extern "C"` entry points; plain pointers and sizes, no C++/torch types).
 * Every entry point replaces one function of the reference's hot path
 * (yuansliu/minicom, paths relative to /root/reference/src):
 *
 *   mcb_for_reads*      <- kt_for_reads        kthread_reads.c:247  (process_reads :40, mm_sketch_two sketch.c:238)
 *   mcb_for_bucket      <- kt_for_bucket       kthread_bucket.c:562 (process_bucket :381, construct_ref :69,
 *                                                                    mm_sketch_lh_ori sketch.c:116)
 *   mcb_idx_build       <- mm_idx_generation   kthread_idx.c:170    (worker_post :116, radix_sort_128x ksort.h:108)
 *   mcb_idx_get         <- mm_idx_get          kthread_idx.c:84
 *   mcb_combine         <- combine_cluster     kthread_cb.c:570     (find_next :220, match_pro :36, construct_ref2 :105, cp_cluster :374;
 *                                               the caller between kt_for_bucket and realign_hash: SURVEY.md 8f, N1)
 *   mcb_realign         <- realign_hash        kthread_hash_realign.c:569 (singleRead2bitset bbhashdict.c:127,
 *                                               constructdictionary_realign :3, realign_hash_search :316)
 *   mcb_readset_*       <- bseq_open/bseq_read bseq.c:19-96  (FASTQ -> 2-bit packed rows on the host)
 *   mcb_dump_encode     <- print_encode        kthread_dump.c:33 (the per-read diff encoding only; SURVEY.md 8f, N2 first slice)
 *   mcb_sketch_lh_host  <- mm_sketch_lh_ori    sketch.c:116  (per-contig call made by the host contig merger,
 *                                               kthread_cb.c:234,365,418 — a boundary helper, not the batched path)
 *
 * All functions return 0 on success and a negative MCB_E* code on failure;
 * mcb_last_error() returns a human-readable message.  There is no CPU
 * fallback: without a usable CUDA device mcb_create() fails.
 *
 * Result structs hold pointers into context-owned pinned host memory that
 * stay valid until the next call of the same entry point or mcb_destroy().
 */
#ifndef MINICOM_B200_H
#define MINICOM_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MCB_OK            0
#define MCB_ECUDA        -1   /* CUDA runtime error (message has the call site) */
#define MCB_EINVAL       -2   /* bad argument / unsupported option value */
#define MCB_ESTATE       -3   /* call order violated (e.g. mcb_for_bucket before mcb_for_reads) */
#define MCB_EINPUT       -4   /* input data outside the supported domain (see message) */
#define MCB_ENOMEM       -5

/* The reference's unit of exchange, mm128_t (breads.h:15-17):
 *   x = hash64(canonical k-mer), y = id<<32 | last_pos<<1 | strand. */
typedef struct { uint64_t x, y; } mcb_tuple;

/* Run-time copy of the reference's compile-time options (minicommain.c:81-143,
 * preprocess.c:87-107).  All fields are the *resolved* values (defaults applied
 * by the caller, see mcb_resolve_params). */
typedef struct {
	int32_t readlen;         /* L, 1..256                      (config.h readlen; minicom:51) */
	int32_t k;               /* -k, <= 31                      (minicommain.c:92-98) */
	int32_t b;               /* bucket bits, 14                (minicommain.c:175) */
	int32_t rw;              /* -w contig minimizer window     (preprocess.c:89-107) */
	int32_t first_mininum;   /* -m minimizers indexed / contig (minicommain.c:117-119) */
	int32_t diff_threshold;  /* -e                             (minicommain.c:100-102) */
	int32_t max_rounds;      /* hidden -R, default 35          (minicommain.c:64,127) */
	int32_t device;          /* CUDA device ordinal */
} mcb_params;

/* Fill *p from the reference's raw option macros (0 = "not given"), exactly as
 * main()/pre_process() derive them: k=31 (17 if L<80); e=4; m=6; max_rounds=35
 * (lowered only if 0<inimaxrounds<35); rw = L/2-k (L>=70) else 3, overridden by iniw>0. */
void mcb_resolve_params(mcb_params *p, int readlen, int inik, int inithr, int iniw, int inim, int inimaxrounds);

typedef struct mcb_ctx mcb_ctx;

int  mcb_create(const mcb_params *p, mcb_ctx **out);
void mcb_destroy(mcb_ctx *ctx);
const char *mcb_last_error(void);
const char *mcb_version(void);

/* ------------------------------------------------------------------ */
/* kt_for_reads (kthread_reads.c:247)                                   */
/* ------------------------------------------------------------------ */

/* Read classes in the order process_reads() tests them (kthread_reads.c:84-226). */
enum {
	MCB_CLS_SKETCHED = 0,  /* N-replaced (if any N), sketched, goes to the buckets   :182-218 */
	MCB_CLS_ALLA     = 1,  /* sp->allA_id   :84  */
	MCB_CLS_ALLT     = 2,  /* sp->allT_id   :95  */
	MCB_CLS_ALLN     = 3,  /* sp->allN_id   :106 */
	MCB_CLS_FPA      = 4,  /* reads->fpA_id :113 */
	MCB_CLS_FPT      = 5,  /* reads->fpT_id :118 */
	MCB_CLS_FPN      = 6,  /* reads->fpN_id :123 */
	MCB_CLS_NFILE    = 7   /* reads->Nfile_id :219 (more than 0.4*L N) */
};

typedef struct {
	uint64_t n_reads;
	const uint8_t *cls;          /* [n_reads] MCB_CLS_* */
	/* reads that contain at least one 'N', ascending rid (seq->n_pos, kthread_reads.c:69-80) */
	uint64_t n_nreads;
	const uint32_t *nread_rid;   /* [n_nreads] */
	const uint8_t  *nread_repl;  /* [n_nreads] replacement base 'A'/'C'/'G'/'T' for sketched reads, 0 otherwise (:185-204) */
	const uint64_t *nread_off;   /* [n_nreads+1] offsets into npos */
	const uint32_t *npos;        /* N positions, ascending within a read */
	uint64_t n_sketched;         /* reads that produced a tuple */
} mcb_reads_result;

/* rows: host memory, n*L bytes, read i at rows+i*L (no terminators), characters in {A,C,G,T,N}. */
int mcb_for_reads(mcb_ctx *ctx, const char *rows, uint64_t n, mcb_reads_result *res);
/* Same, gathering from scattered NUL-terminated strings: string i is *(const char**)((const char*)first_seq_ptr + i*stride)
 * (for the reference: &reads->seq[0].seq with stride sizeof(bseq1_t), bseq.h:10-16).  n_threads host threads gather. */
int mcb_for_reads_ptrs(mcb_ctx *ctx, const void *first_seq_ptr, size_t stride, uint64_t n, int n_threads, mcb_reads_result *res);
/* Same with the rows already resident in device memory (d_rows: device pointer, n*L bytes). */
int mcb_for_reads_device(mcb_ctx *ctx, const char *d_rows, uint64_t n, mcb_reads_result *res);

/* ------------------------------------------------------------------ */
/* FASTQ -> packed reads (bseq.c:19-96), SURVEY.md 8f row N3            */
/* ------------------------------------------------------------------ */
/* A read set on the host in the layout the device works on: row i = row_words 64-bit words, base j of the read in word j/32 at
 * bits 2*(j%32), codes A0 C1 G2 T3 (seq_nt4_table, sketch.c:8-25), unused bits zero; an 'N' is stored as code 0 and recorded
 * in the side table of the reads that contain one: nread_rid ascending, nmask[j*row_words + w] has bit 2*(i%32) of word i/32
 * set where read nread_rid[j] has an N at position i (what process_reads keeps as seq->n_pos, kthread_reads.c:69-80).
 * The rows live in page-locked memory when a CUDA device is present.  mcb_readset_add_fastq replaces bseq_open + bseq_read +
 * bseq_close (a second call appends, like bseq_read_second for -1/-2): the file (plain or gzip, through zlib like the
 * reference) is parsed with kseq_read's record grammar (kseq.h:185-224: FASTA and FASTQ, multi-line sequences, CR LF), every
 * sequence must be `readlen` long (bseq.c:54-57) and consist of A,C,G,T,N (MCB_EINPUT otherwise).  n_threads host threads pack.
 * ascii_out (may be NULL): receives a malloc'd block of n_added rows of readlen+1 bytes, NUL-terminated — the strings the
 * reference's host stages keep in reads->seq[i].seq; the caller frees it. */
typedef struct mcb_readset mcb_readset;
typedef struct {
	uint64_t n_reads;
	int32_t readlen, row_words;
	const uint64_t *packed;      /* [n_reads][row_words] */
	uint64_t n_nreads;
	const uint32_t *nread_rid;   /* [n_nreads] */
	const uint64_t *nmask;       /* [n_nreads][row_words] */
} mcb_readset_view;
int  mcb_readset_create(int readlen, mcb_readset **out);
void mcb_readset_destroy(mcb_readset *rs);
int  mcb_readset_add_fastq(mcb_readset *rs, const char *path, int n_threads, char **ascii_out, uint64_t *n_added);
int  mcb_readset_add_fastq_buffer(mcb_readset *rs, const char *buf, uint64_t len, int n_threads, char **ascii_out, uint64_t *n_added);
/* the same packing for reads that are already rows of readlen characters in memory */
int  mcb_readset_add_rows(mcb_readset *rs, const char *rows, uint64_t n, int n_threads);
void mcb_readset_get(const mcb_readset *rs, mcb_readset_view *view);   /* valid until the next add / destroy */

/* kt_for_reads on packed reads: same results as mcb_for_reads on the ASCII rows they were packed from, ceil(L/4) (rounded to
 * 16) bytes per read over PCIe instead of L.  packed / nread_rid / nmask: host memory in the mcb_readset_view layout
 * (nread_rid relative to this call's first row).  Rows with non-zero unused bits, masks outside the read or over a non-zero
 * code are rejected (MCB_EINPUT). */
int mcb_for_reads_packed(mcb_ctx *ctx, const uint64_t *packed, uint64_t n, const uint32_t *nread_rid, const uint64_t *nmask, uint64_t n_nreads,
                         mcb_reads_result *res);
/* Same with the three arrays already resident in device memory (16-byte aligned). */
int mcb_for_reads_packed_device(mcb_ctx *ctx, const uint64_t *d_packed, uint64_t n, const uint32_t *d_nread_rid, const uint64_t *d_nmask, uint64_t n_nreads,
                                mcb_reads_result *res);

/* Test/debug view of B[0] (kthread_reads.c:213-218): one tuple per read in rid order,
 * {~0,~0} for reads that were not sketched.  out: host, n_reads tuples. */
int mcb_debug_read_tuples(mcb_ctx *ctx, mcb_tuple *out);
/* Batched mm_sketch_two (sketch.c:238) over reads already loaded, with an arbitrary k (used by rounds >= 2). */
int mcb_debug_sketch_two(mcb_ctx *ctx, const uint32_t *rids, uint64_t n, int k, mcb_tuple *out);
/* Copy the 2-bit packed, N-replaced reads back as ASCII rows (n_reads*L bytes). */
int mcb_debug_unpack_reads(mcb_ctx *ctx, char *rows_out);

/* ------------------------------------------------------------------ */
/* kt_for_bucket (kthread_bucket.c:562)                                 */
/* ------------------------------------------------------------------ */
typedef struct {
	/* seed contigs in the order the single-threaded reference creates them:
	 * round, bucket, minimizer ascending (reads->clusters[0][0]) */
	uint64_t n_clusters;
	const uint32_t *cl_n;        /* [n_clusters] members per cluster (cluster_t.n) */
	const uint64_t *cl_a_off;    /* [n_clusters+1] */
	const uint64_t *cl_a;        /* members: rid<<32 | offset<<1 | dir (cluster_t.a, kthread_bucket.c:349) */
	const uint64_t *cl_ref_off;  /* [n_clusters+1] */
	const char     *cl_ref;      /* consensus strings, concatenated, no terminators (cluster_t.ref) */
	/* reads->sg in push order (kthread_bucket.c:198-203,402-413,482-485) */
	uint64_t n_sg;
	const uint32_t *sg;
	/* tuples pushed into reads->mi[0] (kthread_bucket.c:458-474): for cluster c, tuples mi[c*m .. c*m+mi_cnt[c]) */
	const uint8_t  *mi_cnt;      /* [n_clusters] <= first_mininum */
	const mcb_tuple *mi;         /* [n_clusters*first_mininum] */
	int32_t rounds;              /* rounds executed (kmer = k-1 .. k-rounds) */
	uint64_t n_sketched_total;   /* tuples sketched over all rounds (N_sk of the roofline formula) */
	uint64_t n_grouped;          /* reads that sat in a group of >=2 (N_grp) */
} mcb_bucket_result;

int mcb_for_bucket(mcb_ctx *ctx, mcb_bucket_result *res);
/* Same for a caller that continues with mcb_combine on this context: the seed contigs and their index tuples stay on the device;
 * only n_clusters, the singles (n_sg, sg) and the counters of *res are filled, the other pointers are NULL. */
int mcb_for_bucket_keep(mcb_ctx *ctx, mcb_bucket_result *res);

/* kt_for_bucket's loop control (kthread_bucket.c:584-585,594,607-622), for callers that drive the rounds themselves:
 *   is_last = mcb_round_control_begin(&rc, k, max_rounds);  ... one round ...;  stop = mcb_round_control_end(&rc, members so far) */
typedef struct { int32_t round, last_rounds; int64_t pre_members; } mcb_round_control;
void mcb_round_control_init(mcb_round_control *rc);
int  mcb_round_control_begin(mcb_round_control *rc, int k, int max_rounds);
int  mcb_round_control_end(mcb_round_control *rc, uint64_t members_total);

/* ------------------------------------------------------------------ */
/* mm_idx_generation / mm_idx_get (kthread_idx.c:170,84)                */
/* ------------------------------------------------------------------ */
typedef struct mcb_index mcb_index;

/* tuples: host, bucket-major (bucket = x & (2^b-1)), inside a bucket in push order;
 * bucket_off: [2^b+1].  The index reproduces the reference's posting order
 * (radix_sort_128x is unstable above 64 elements; ksort.h:108-157). */
int mcb_idx_build(mcb_ctx *ctx, const mcb_tuple *tuples, const uint64_t *bucket_off, mcb_index **out);
/* Same as mcb_idx_build but the per-bucket arrays are still scattered: bucket i has
 * cnt[i] tuples at ptrs[i] (the reference's mi->B[i].a.{a,n}); n_threads host threads gather them (mm_idx_generation's n_threads). */
int mcb_idx_build_scattered(mcb_ctx *ctx, const mcb_tuple *const *ptrs, const uint64_t *cnt, int n_threads, mcb_index **out);
/* Host lookup; thread-safe; returns pointer to *n postings (y values) or NULL. */
const uint64_t *mcb_idx_get(const mcb_index *idx, uint64_t minier, int *n);
void mcb_idx_destroy(mcb_index *idx);
/* Introspection for tests: number of distinct keys and postings. */
void mcb_idx_stats(const mcb_index *idx, uint64_t *n_keys, uint64_t *n_post);
/* The whole index as flat host arrays (valid until mcb_idx_destroy): keys[n_keys] distinct minimizers, bucket-major and ascending
 * inside a bucket; kstart[n_keys+1] posting offsets; post[n_post] the y values, per key in the reference's order;
 * bucket_keys[2^b+1] first key of every bucket.  Any of the four pointers may be NULL. */
void mcb_idx_arrays(const mcb_index *idx, const uint64_t **keys, const uint32_t **kstart, const uint64_t **post, const uint32_t **bucket_keys);

/* ------------------------------------------------------------------ */
/* combine_cluster (kthread_cb.c:570): the contig merge                 */
/* ------------------------------------------------------------------ */
typedef struct {
	/* the contigs after the merge, in the order the single-threaded reference leaves them in reads->clusters[idxv][0]:
	 * per iteration the merged contigs in the order of their first partner, then the untouched ones (kthread_cb.c:460-494) */
	uint64_t n_clusters;
	const uint32_t *cl_n;        /* [n_clusters] members per contig */
	const uint64_t *cl_a_off;    /* [n_clusters+1] */
	const uint64_t *cl_a;        /* members: rid<<32 | offset<<1 | dir; merged contigs sorted by (offset, dir) (construct_ref2, :107) */
	const uint64_t *cl_ref_off;  /* [n_clusters+1] */
	const char     *cl_ref;      /* consensus strings, concatenated, no terminators */
	int32_t iterations;          /* iterations of the loop at :573-627 = mm_idx_generation calls made (idxv = iterations & 1) */
	uint64_t n_merges;           /* pairs merged over all iterations */
	uint64_t n_index_tuples;     /* tuples indexed over all iterations (T_cb of the roofline formula) */
} mcb_combine_result;

/* Runs on the seed contigs mcb_for_bucket left on the device (call it right after mcb_for_bucket, same context): per iteration
 * the minimizer index of the contigs (device-resident, reference posting order), every contig's (w,k)-minimizers, the ordered
 * lists of partners that pass the strand and match_pro <= cbthreshold tests, the first-come resolution in contig order, and the
 * merged consensus strings.  cbthreshold: the reference's global (2 * diff_threshold unless -g, minicommain.c:122-126).
 * The index lookups (mm_idx_get) and the sketches of kthread_cb.c happen on the device; nothing is uploaded. */
int mcb_combine(mcb_ctx *ctx, int cbthreshold, mcb_combine_result *res);

/* ------------------------------------------------------------------ */
/* realign_hash (kthread_hash_realign.c:569)                            */
/* ------------------------------------------------------------------ */
typedef struct {
	/* reads claimed by contigs, in the exact order the single-threaded reference appends them
	 * (contig, window jj, forward-then-reverse, dictionary l ascending, sg index descending) */
	uint64_t n_claims;
	const uint32_t *claim_contig; /* index into the contig list passed in */
	const uint32_t *claim_sg;     /* index into sg (sg_flag[claim_sg]=true) */
	const uint64_t *claim_y;      /* rid<<32 | jj<<1 | dir  (kthread_hash_realign.c:405,472) */
	const uint64_t *claim_prio;   /* (window index over all contigs)<<5 | reverse<<4 | dictionary: the step of the reference's
	                                 window loop that made the claim; ascending along the list (sharding: the merge key) */
	/* singles diverted to the near-poly-A / near-poly-T lists (bbhashdict.c:157-216), ascending sg index */
	uint64_t n_fpA, n_fpT;
	const uint32_t *fpA_sg, *fpT_sg;
	/* work counters for the roofline formula: contig windows; dictionary probes the reference's window loop
	 * would issue for them (kthread_hash_realign.c:355-504); (window, single) pairs with equal dictionary keys that were
	 * verified; n_dict_keys is reserved (0) */
	uint64_t n_windows, n_probes, n_candidates, n_dict_keys;
	int32_t numdict;              /* numdict_s actually used */
} mcb_realign_result;

/* sg: host, n_sg read ids (reads->sg after updateSingle); refs/ref_off: host, contig consensus strings
 * concatenated (ref_off[n_contigs+1]); threshold: the current step of the -e/-S/-E schedule;
 * maxsearch: the reference's global (500, or 2000 when sg.n<=5M, preprocess.c:169-172);
 * ininumdict: raw -s value (0 = not given), resolved as setglobalarrays_realign does (:150-171).
 * The contigs do not change between the rounds of one -e/-S/-E schedule (preprocess.c:197-232: only updateSingle() runs
 * between two realign_hash calls), and the library keeps the device-side k-mer table it built over them: passing
 * refs == NULL and ref_off == NULL reuses the contigs of the previous call (n_contigs must be 0 or the same count).
 * When refs is given, the contigs are re-packed only if the strings differ from the cached ones.  The k-mer table holds
 * the contig k-mers that the singles of the call that built it can ask for; a later call reuses it when its singles are
 * a subset of those (what updateSingle() guarantees) and rebuilds it transparently otherwise.
 * Limits of one call: fewer than 2^32 contig bases and fewer than 2^31 (single, dictionary) pairs (MCB_EINVAL beyond). */
int mcb_realign(mcb_ctx *ctx, const uint32_t *sg, uint64_t n_sg, const char *refs, const uint64_t *ref_off,
                uint64_t n_contigs, int threshold, int maxsearch, int ininumdict, mcb_realign_result *res);

#define MCB_CLAIM_NONE 0x7F7F7F7F7F7F7F7Fll

/* ------------------------------------------------------------------ */
/* print_encode (kthread_dump.c:33-236), SURVEY.md 8f row N2: first slice */
/* ------------------------------------------------------------------ */
/* The per-read diff encoding of the dump stage (the loop at kthread_dump.c:66-118, shared by the ORDER, default and _PE
 * variants): for every member of every contig the line print_encode writes to dif_char.txt — the read as it was before N
 * replacement, reverse-complemented when dir, against its consensus window: runs of >= 2 equal characters as their decimal
 * length, a run of 1 copied, every mismatching character copied, the trailing run dropped, "0" when nothing differs.
 * members: y = rid<<32 | pos<<1 | dir, contig-major in the order the caller dumps them (after its qsort by cmpcluster2/3);
 * member_off[n_contigs+1]; refs / ref_off as in mcb_realign (NULL, NULL: the contigs the context holds from the last
 * mcb_combine / mcb_realign).  The reads are the ones mcb_for_reads* loaded.  enc holds the encodings back to back, no
 * separators; member i occupies [enc_off[i], enc_off[i+1]).  Not yet bound by the drop-in shim: the rest of cluster_dump
 * (position deltas, direction bits, packed consensus, the output files) is still the reference's host code. */
typedef struct {
	uint64_t n_members, n_bytes;
	const uint64_t *enc_off;      /* [n_members+1] */
	const char *enc;              /* [n_bytes] */
} mcb_encode_result;
int mcb_dump_encode(mcb_ctx *ctx, const uint64_t *members, const uint64_t *member_off, const char *refs, const uint64_t *ref_off, uint64_t n_contigs,
                    mcb_encode_result *res);

/* ------------------------------------------------------------------ */
/* the same path over the GPUs of one box                */
/* ------------------------------------------------------------------ */
/* One context per GPU — one per process (torchrun-style launch) or several in one process, one host thread each (the drop-in
 * host program: dropin/mcb_dropin.cpp with MCB_DEVICES=0,1,...).  The contexts of a job share an NCCL communicator; all data
 * movement between GPUs (grouped ncclSend/ncclRecv over NVLink) happens inside the library, on the context's stream.
 *
 *   reads    contiguous read-id ranges: rank r loads [rid_base, rid_base + n) with mcb_for_reads* after mcb_shard_begin
 *   buckets  the 16384 minimizer buckets in contiguous ranges, owner = bucket * n_ranks >> 14; every round of kt_for_bucket
 *            sends each tuple, together with the 2-bit packed row of its read, to the owner of its bucket
 *   index    every rank builds (mcb_idx_build) the buckets it owns
 *   Stage 2  a rank realigns the singles it produced in Stage 1 against ALL contigs (mcb_shard_realign)
 *
 * The job's results are the ranks' results concatenated in rank order — round by round for kt_for_bucket (round_counts) — and
 * the claim lists merged by (claim_prio ascending, claim_sg descending); that reproduces one GPU (= the single-threaded
 * reference) bit for bit.  Functions marked "collective" must be called by every rank of the communicator. */
#define MCB_NCCL_ID_BYTES 128
int mcb_shard_unique_id(void *id128);                                               /* ncclGetUniqueId; one rank makes it, all ranks pass it to mcb_shard_init */
int mcb_shard_init(mcb_ctx *ctx, const void *id128, int rank, int n_ranks);        /* collective: ncclCommInitRank on the context's device */
int mcb_shard_attach(mcb_ctx *ctx, void *nccl_comm, int rank, int n_ranks);        /* or: use a communicator (ncclComm_t) the caller made, e.g. with ncclCommInitAll */
int mcb_shard_begin(mcb_ctx *ctx, uint64_t n_total, uint64_t rid_base);            /* a new job of n_total reads; read ids are global from here on */
/* kt_for_bucket over the job (collective).  res: this rank's seed contigs / singles / index tuples in its own round order, contig
 * ids inside res->mi global; round_counts[4 r .. 4 r + 3] = contigs, members, consensus bytes, singles contributed by round r. */
int mcb_shard_for_bucket(mcb_ctx *ctx, mcb_bucket_result *res, uint64_t *round_counts, int cap_rounds);
/* realign_hash over the job (collective).  sg: the singles of the job's list that THIS rank produced in Stage 1, in the list's
 * order; sg_index: their positions in the job's list (ascending); n_sg_total: length of the job's list.  refs / ref_off: ALL
 * contigs (NULL, NULL reuses the previous call's, as in mcb_realign).  res: the claims on this rank's singles in the reference's
 * append order, claim_contig global, claim_sg / fpA_sg / fpT_sg positions in the job's list.  Dictionary bins span the singles
 * of the whole job: their sizes are bounded collectively, and bins above maxsearch are replayed like on one GPU. */
int mcb_shard_realign(mcb_ctx *ctx, const uint32_t *sg, const uint32_t *sg_index, uint64_t n_sg_local, uint64_t n_sg_total,
                      const char *refs, const uint64_t *ref_off, uint64_t n_contigs, int threshold, int maxsearch, int ininumdict, mcb_realign_result *res);

/* Several GPUs behind the single-GPU call shapes, for a host program that is one process (what the drop-in shim uses when
 * MCB_DEVICES lists more than one device): one context per device, one host thread per context while a call runs, results
 * merged into exactly what one GPU returns.  The arguments and result structs mean what they mean in mcb_for_reads_ptrs,
 * mcb_for_bucket, mcb_idx_build_scattered (the returned index answers mcb_idx_get / mcb_idx_destroy) and mcb_realign. */
typedef struct mcb_group mcb_group;
int  mcb_group_create(const mcb_params *p, const int *devices, int n_devices, mcb_group **out);   /* p->device is ignored */
void mcb_group_destroy(mcb_group *g);
int  mcb_group_size(const mcb_group *g);
mcb_ctx *mcb_group_context(mcb_group *g, int rank);                                                 /* e.g. for mcb_timers_enable */
int  mcb_group_for_reads_ptrs(mcb_group *g, const void *first_seq_ptr, size_t stride, uint64_t n, int n_threads, mcb_reads_result *res);
int  mcb_group_for_bucket(mcb_group *g, mcb_bucket_result *res);
int  mcb_group_idx_build_scattered(mcb_group *g, const mcb_tuple *const *ptrs, const uint64_t *cnt, int n_threads, mcb_index **out);
int  mcb_group_realign(mcb_group *g, const uint32_t *sg, uint64_t n_sg, const char *refs, const uint64_t *ref_off, uint64_t n_contigs,
                       int threshold, int maxsearch, int ininumdict, mcb_realign_result *res);
size_t mcb_group_timers_dump(mcb_group *g, char *buf, size_t cap);                                  /* per timer: the maximum over the ranks */

/* ------------------------------------------------------------------ */
/* host-side boundary helpers                                           */
/* ------------------------------------------------------------------ */
/* mm_sketch_lh_ori (sketch.c:116): windowed (w,k)-minimizers of one string.  Writes at most cap tuples to out,
 * returns the number the reference would have produced (may exceed cap). */
int64_t mcb_sketch_lh_host(const char *str, int len, int w, int k, uint32_t rid, mcb_tuple *out, int64_t cap);
/* mm_sketch_two (sketch.c:238) for one string (host; used by tests and by INTEGRATION examples). */
void mcb_sketch_two_host(const char *str, int len, int k, uint32_t rid, mcb_tuple *out);
/* hash64 (sketch.c:27) */
uint64_t mcb_hash64(uint64_t key, uint64_t mask);

/* ------------------------------------------------------------------ */
/* measurement                                                          */
/* ------------------------------------------------------------------ */
/* Device-time accounting (CUDA events on the library's stream).  Timer names: "for_reads", "for_bucket",
 * "idx_build", "realign" (whole entry point, device work only), "h2d", "d2h", per-kernel names "k:<kernel>", and, sharded,
 * "nccl:<what>" (CUDA-event time of the collectives) plus "nccl_bytes_sent" (bytes, not milliseconds).
 * mcb_timer_get returns accumulated milliseconds and the launch count; mcb_timers_reset zeroes everything. */
void   mcb_timers_enable(mcb_ctx *ctx, int on);
void   mcb_timers_reset(mcb_ctx *ctx);
double mcb_timer_get(mcb_ctx *ctx, const char *name, uint64_t *count);
/* Writes "name ms count\n" lines for every timer into buf (NUL-terminated, truncated to cap); returns needed size. */
size_t mcb_timers_dump(mcb_ctx *ctx, char *buf, size_t cap);
/* Total kernels launched by this context since creation / last reset. */
uint64_t mcb_kernel_launches(mcb_ctx *ctx);

#ifdef __cplusplus
}
#endif
#endif /* MINICOM_B200_H */
