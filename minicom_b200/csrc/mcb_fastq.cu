// mcb_fastq.cu — N3: FASTQ -> 2-bit packed reads on the host.
//
// Replaces bseq_open / bseq_read / bseq_read_second / bseq_close (bseq.c:19-96: kseq records, one strdup per read) and the
// ASCII half of process_reads (kthread_reads.c:56-80: per-character counting, N positions): the file is parsed once, every
// read goes straight into the layout the device works on (u64[n][WS], 2 bits per base, N as code 0 plus a side table of
// N masks for the few reads that contain one), in page-locked memory, so kt_for_reads uploads ceil(L/4) bytes per read
// instead of L.  Host code only: no kernel here (the device half is mcb_for_reads_packed, mcb_stage1.cu).
//
// Record grammar = kseq_read (kseq.h:185-224), restated on an in-memory buffer:
//   * the next '>' or '@' anywhere starts a header (only when the previous record ended on its quality string);
//   * name up to the first white space, the rest of the line is a comment;
//   * sequence = every following line, empty lines skipped, until a line that starts with '>', '+' or '@';
//   * '+' line skipped; quality lines are appended until they are at least as long as the sequence; a record whose quality is
//     missing or of another length ends the file (kseq_read returns -2 and bseq_read's loop stops, bseq.c:44);
//   * a trailing '\r' is dropped from a line when the string so far is longer than one character (kseq.h:138).
// Every sequence must be `readlen` long (bseq.c:54-57: the reference prints and exits; here MCB_EINPUT).
#include "mcb_common.cuh"
#include <zlib.h>
#include <stdlib.h>
#include <algorithm>
#include <thread>
#include <atomic>

struct mcb_readset {
	int L = 0, Wd = 0, WS = 0;
	uint64_t n = 0, cap = 0;          // reads held / rows allocated
	uint64_t *packed = nullptr;
	bool pinned = false;
	std::vector<uint32_t> nrid;       // reads with N, ascending
	std::vector<uint64_t> nmask;      // [nrid.size()][WS]
};

static void rs_free_rows(mcb_readset *rs)
{
	if (!rs->packed) return;
	if (rs->pinned) cudaFreeHost(rs->packed); else free(rs->packed);
	rs->packed = nullptr; rs->cap = 0;
}

// page-locked when a CUDA device is there (the upload is then a plain DMA); ordinary memory otherwise, so that the parser can
// be exercised where there is no GPU (mcb_for_reads_packed stages pageable rows like every other entry point)
static int rs_reserve(mcb_readset *rs, uint64_t rows)
{
	if (rows <= rs->cap) return MCB_OK;
	uint64_t want = std::max<uint64_t>(rows, rs->cap + rs->cap / 2);
	size_t bytes = (size_t)want * rs->WS * 8 + 64;
	uint64_t *p = nullptr; bool pin = true;
	if (cudaMallocHost((void**)&p, bytes) != cudaSuccess) {
		cudaGetLastError(); pin = false;
		if (posix_memalign((void**)&p, 64, bytes)) { mcb_set_error("out of memory for %llu packed reads", (unsigned long long)want); return MCB_ENOMEM; }
	}
	if (rs->n) memcpy(p, rs->packed, (size_t)rs->n * rs->WS * 8);
	rs_free_rows(rs);
	rs->packed = p; rs->pinned = pin; rs->cap = want;
	return MCB_OK;
}

extern "C" int mcb_readset_create(int readlen, mcb_readset **out)
{
	if (!out) { mcb_set_error("mcb_readset_create: null argument"); return MCB_EINVAL; }
	*out = nullptr;
	if (readlen < 12 || readlen > 256) { mcb_set_error("readlen %d out of range [12,256] (minicom:51-54)", readlen); return MCB_EINVAL; }
	mcb_readset *rs = new mcb_readset();
	rs->L = readlen; rs->Wd = (readlen + 31) / 32; rs->WS = (rs->Wd + 1) & ~1;
	*out = rs;
	return MCB_OK;
}

extern "C" void mcb_readset_destroy(mcb_readset *rs)
{
	if (!rs) return;
	rs_free_rows(rs);
	delete rs;
}

extern "C" void mcb_readset_get(const mcb_readset *rs, mcb_readset_view *v)
{
	v->n_reads = rs->n; v->readlen = rs->L; v->row_words = rs->WS; v->packed = rs->packed;
	v->n_nreads = rs->nrid.size(); v->nread_rid = rs->nrid.data(); v->nmask = rs->nmask.data();
}

// ---------------------------------------------------------------- ASCII -> 2 bits
// seq_nt4_table (sketch.c:8-25) narrowed to what process_reads accepts: upper-case A,C,G,T -> 0..3, 'N' -> 4; everything else 5
// (rejected: the reference's behaviour on it is undefined, see mcb_code_of)
static const uint8_t *code_table()
{
	static uint8_t t[256];
	static std::atomic<int> ready(0);
	if (!ready.load(std::memory_order_acquire)) {
		uint8_t u[256];
		memset(u, 5, sizeof u);
		u[(unsigned char)'A'] = 0; u[(unsigned char)'C'] = 1; u[(unsigned char)'G'] = 2; u[(unsigned char)'T'] = 3; u[(unsigned char)'N'] = 4;
		memcpy(t, u, sizeof u);
		ready.store(1, std::memory_order_release);
	}
	return t;
}

// one read: L characters -> WS words (+ N mask); returns 0 ok, 1 has N, -1 bad character
static inline int pack_row(const uint8_t *tab, const unsigned char *s, int L, int Wd, int WS, uint64_t *row, uint64_t *mask)
{
	unsigned bad = 0;
	uint64_t anyn = 0;
	for (int w = 0; w < Wd; ++w) {
		const int lim = std::min(32, L - w * 32);
		const unsigned char *p = s + w * 32;
		uint64_t word = 0, nm = 0;
		for (int j = 0; j < lim; ++j) {
			const unsigned c = tab[p[j]];
			bad |= c > 4u;
			word |= (uint64_t)(c < 4u ? c : 0u) << (2 * j);
			nm |= (uint64_t)(c == 4u) << (2 * j);
		}
		row[w] = word; mask[w] = nm; anyn |= nm;
	}
	for (int w = Wd; w < WS; ++w) { row[w] = 0; mask[w] = 0; }
	if (bad) return -1;
	return anyn ? 1 : 0;
}

struct RowSource {                   // where the characters of read i are
	const char *rows = nullptr;      // contiguous rows of L characters, or
	const char *buf = nullptr;       // a parsed file: seq_off[i] into buf, or (top bit set) into arena
	const uint64_t *seq_off = nullptr;
	const char *arena = nullptr;
	int L = 0;
	const unsigned char *at(uint64_t i) const
	{
		if (rows) return (const unsigned char*)rows + i * (size_t)L;
		const uint64_t o = seq_off[i];
		return (const unsigned char*)((o >> 63) ? arena + (o & ~(1ull << 63)) : buf + o);
	}
};

// packs reads [0, n) of src behind the reads already in rs; ascii_out (optional): n rows of L+1 bytes, NUL-terminated
static int rs_append(mcb_readset *rs, const RowSource &src, uint64_t n, int n_threads, char *ascii_out)
{
	if (rs->n + n >= (1ull << 31)) { mcb_set_error("too many reads (rid is a signed 32-bit int in the reference, kthread_bucket.c:48)"); return MCB_EINVAL; }
	MCB_TRY(rs_reserve(rs, rs->n + n));
	const int L = rs->L, Wd = rs->Wd, WS = rs->WS;
	const uint8_t *tab = code_table();
	int T = (int)std::min<uint64_t>((uint64_t)std::max(1, n_threads), (n + 16383) / 16384);
	if (T < 1) T = 1;
	std::vector<std::vector<uint32_t>> t_rid((size_t)T);
	std::vector<std::vector<uint64_t>> t_mask((size_t)T);
	std::vector<uint64_t> t_bad((size_t)T, ~0ull);
	uint64_t *base = rs->packed + rs->n * (size_t)WS;
	const uint64_t rid0 = rs->n;
	auto work = [&](int t) {
		const uint64_t a = n * t / T, b = n * (t + 1) / T;
		uint64_t mask[8];
		for (uint64_t i = a; i < b; ++i) {
			const unsigned char *s = src.at(i);
			const int r = pack_row(tab, s, L, Wd, WS, base + i * (size_t)WS, mask);
			if (r < 0) { if (t_bad[t] == ~0ull) t_bad[t] = i; continue; }
			if (r > 0) { t_rid[t].push_back((uint32_t)(rid0 + i)); t_mask[t].insert(t_mask[t].end(), mask, mask + WS); }
			if (ascii_out) { char *d = ascii_out + i * (size_t)(L + 1); memcpy(d, s, L); d[L] = 0; }
		}
	};
	if (T == 1) work(0);
	else {
		std::vector<std::thread> th;
		for (int t = 0; t < T; ++t) th.emplace_back(work, t);
		for (auto &x : th) x.join();
	}
	for (int t = 0; t < T; ++t)
		if (t_bad[t] != ~0ull) {
			mcb_set_error("read %llu contains characters other than A,C,G,T,N (unsupported; the reference's behaviour on them is undefined)", (unsigned long long)(rid0 + t_bad[t]));
			return MCB_EINPUT;
		}
	for (int t = 0; t < T; ++t) {
		rs->nrid.insert(rs->nrid.end(), t_rid[t].begin(), t_rid[t].end());
		rs->nmask.insert(rs->nmask.end(), t_mask[t].begin(), t_mask[t].end());
	}
	rs->n += n;
	return MCB_OK;
}

extern "C" int mcb_readset_add_rows(mcb_readset *rs, const char *rows, uint64_t n, int n_threads)
{
	if (!rs || (n && !rows)) { mcb_set_error("mcb_readset_add_rows: null argument"); return MCB_EINVAL; }
	RowSource src; src.rows = rows; src.L = rs->L;
	return rs_append(rs, src, n, n_threads, nullptr);
}

// ---------------------------------------------------------------- kseq_read on a buffer
static inline bool ks_isspace(unsigned char c) { return c == ' ' || (c >= '\t' && c <= '\r'); }

// Appends the sequence offsets of every record of buf[0,len) to seq_off; sequences that span several lines are joined in arena.
// Returns MCB_OK, or MCB_EINPUT when a sequence is not L long.
static int index_records(const char *buf, size_t len, int L, std::vector<uint64_t> &seq_off, std::string &arena, uint64_t first_rid)
{
	size_t pos = 0;
	int last_char = 0;
	std::string joined;
	for (;;) {
		if (last_char == 0) {                                    // jump to the next header character (kseq.h:189-193)
			while (pos < len && buf[pos] != '>' && buf[pos] != '@') ++pos;
			if (pos >= len) break;
			last_char = buf[pos++];
		}
		if (pos >= len) break;                                   // ks_getuntil at end of file: kseq_read returns -1 (:195)
		{   // name, then the comment (:195-196)
			size_t i = pos;
			while (i < len && !ks_isspace((unsigned char)buf[i])) ++i;
			if (i < len) {
				const char c = buf[i];
				pos = i + 1;
				if (c != '\n') { const char *nl = (const char*)memchr(buf + pos, '\n', len - pos); pos = nl ? (size_t)(nl - buf) + 1 : len; }
			} else pos = len;
		}
		// sequence lines (:201-205)
		size_t s_off = 0, s_len = 0;
		int n_lines = 0, c = -1;
		for (;;) {
			if (pos >= len) { c = -1; break; }
			c = (unsigned char)buf[pos++];
			if (c == '>' || c == '+' || c == '@') break;
			if (c == '\n') continue;
			const size_t a = pos - 1;
			const char *nl = (const char*)memchr(buf + pos, '\n', len - pos);
			const size_t e = nl ? (size_t)(nl - buf) : len;
			pos = nl ? e + 1 : len;
			if (n_lines == 0) { s_off = a; s_len = e - a; }
			else {
				if (n_lines == 1) joined.assign(buf + s_off, s_len);
				joined.append(buf + a, e - a);
				s_len = joined.size();
			}
			++n_lines;
			if (s_len > 1 && (n_lines == 1 ? buf[s_off + s_len - 1] : joined[s_len - 1]) == '\r') { --s_len; if (n_lines > 1) joined.resize(s_len); }
		}
		if (c == '>' || c == '@') last_char = c;
		if (c == '+') {                                          // FASTQ: skip the '+' line, read the quality (:218-222)
			const char *nl = pos < len ? (const char*)memchr(buf + pos, '\n', len - pos) : nullptr;
			if (!nl) break;                                      // no quality string: -2, the reading loop stops (bseq.c:44)
			pos = (size_t)(nl - buf) + 1;
			size_t ql = 0; char qlast = 0;
			do {
				if (pos >= len) break;
				const char *q = (const char*)memchr(buf + pos, '\n', len - pos);
				const size_t e = q ? (size_t)(q - buf) : len;
				if (e > pos) { ql += e - pos; qlast = buf[e - 1]; }
				pos = q ? e + 1 : len;
				if (ql > 1 && qlast == '\r') { --ql; qlast = 0; }
			} while (ql < s_len);
			last_char = 0;
			if (ql != s_len) break;                              // quality of another length: -2
		}
		if ((int64_t)s_len != (int64_t)L) {
			mcb_set_error("read %llu is %llu characters long, readlen is %d: \"Length of reads are different. The program can not compress it.\" (bseq.c:54-57)",
			              (unsigned long long)(first_rid + seq_off.size()), (unsigned long long)s_len, L);
			return MCB_EINPUT;
		}
		if (n_lines <= 1) seq_off.push_back((uint64_t)s_off);
		else { seq_off.push_back((1ull << 63) | (uint64_t)arena.size()); arena.append(joined.data(), s_len); }
		if (c == -1) break;                                      // end of file inside the sequence: the next call finds nothing
	}
	return MCB_OK;
}

// the whole file in memory, through zlib like the reference (gzopen reads plain files too, bseq.c:23)
static int slurp(const char *path, char **out, size_t *out_len)
{
	gzFile f = gzopen(path, "r");
	if (!f) { mcb_set_error("cannot open %s", path); return MCB_EINPUT; }
	gzbuffer(f, 1u << 20);
	size_t hint = 0;
	if (FILE *fp = fopen(path, "rb")) { if (!fseek(fp, 0, SEEK_END)) { long s = ftell(fp); if (s > 0) hint = (size_t)s; } fclose(fp); }
	size_t cap = std::max<size_t>(hint + 1, 1u << 20), n = 0;
	char *p = (char*)malloc(cap);
	if (!p) { gzclose(f); mcb_set_error("out of memory reading %s", path); return MCB_ENOMEM; }
	for (;;) {
		if (n == cap) {
			cap += cap / 2;
			char *q = (char*)realloc(p, cap);
			if (!q) { free(p); gzclose(f); mcb_set_error("out of memory reading %s", path); return MCB_ENOMEM; }
			p = q;
		}
		const unsigned want = (unsigned)std::min<size_t>(cap - n, 1u << 30);
		const int got = gzread(f, p + n, want);
		if (got < 0) { free(p); gzclose(f); mcb_set_error("read error on %s", path); return MCB_EINPUT; }
		if (got == 0) break;
		n += (size_t)got;
	}
	gzclose(f);
	*out = p; *out_len = n;          // the caller frees it
	return MCB_OK;
}

extern "C" int mcb_readset_add_fastq(mcb_readset *rs, const char *path, int n_threads, char **ascii_out, uint64_t *n_added)
{
	if (!rs || !path) { mcb_set_error("mcb_readset_add_fastq: null argument"); return MCB_EINVAL; }
	if (ascii_out) *ascii_out = nullptr;
	if (n_added) *n_added = 0;
	char *buf = nullptr; size_t len = 0;
	MCB_TRY(slurp(path, &buf, &len));
	const int rc = mcb_readset_add_fastq_buffer(rs, buf, len, n_threads, ascii_out, n_added);
	free(buf);
	return rc;
}

extern "C" int mcb_readset_add_fastq_buffer(mcb_readset *rs, const char *buf, uint64_t len, int n_threads, char **ascii_out, uint64_t *n_added)
{
	if (!rs || (len && !buf)) { mcb_set_error("mcb_readset_add_fastq_buffer: null argument"); return MCB_EINVAL; }
	if (ascii_out) *ascii_out = nullptr;
	if (n_added) *n_added = 0;
	std::vector<uint64_t> seq_off;
	std::string arena;
	seq_off.reserve((size_t)(len / (2 * (size_t)rs->L + 8)) + 16);
	MCB_TRY(index_records(buf, (size_t)len, rs->L, seq_off, arena, rs->n));
	const uint64_t n = seq_off.size();
	char *ascii = nullptr;
	if (ascii_out && n) {
		ascii = (char*)malloc((size_t)n * (rs->L + 1));
		if (!ascii) { mcb_set_error("out of memory for %llu reads", (unsigned long long)n); return MCB_ENOMEM; }
	}
	RowSource src; src.buf = buf; src.seq_off = seq_off.data(); src.arena = arena.data(); src.L = rs->L;
	const int rc = rs_append(rs, src, n, n_threads, ascii);
	if (rc != MCB_OK) { free(ascii); return rc; }
	if (ascii_out) *ascii_out = ascii;
	if (n_added) *n_added = n;
	return MCB_OK;
}
