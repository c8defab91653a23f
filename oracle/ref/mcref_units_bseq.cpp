// TEST INFRASTRUCTURE (oracle/_ref build only; never linked into the product).
// The UNMODIFIED reference's FASTQ reader through ctypes: bseq.c (bseq_open / bseq_read / bseq_close, which instantiates
// kseq.h's kseq_read over gzread) is compiled from /root/reference/src as its own translation unit by build_ref.sh; this file
// only declares and calls it.  seq_nt4_table comes from sketch.c via mcref_units.cpp.
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include "bseq.h"

extern "C" {

// every read of the file, as bseq_read returns them: rows of seq_len characters.  Returns the number of reads (bseq_read itself
// exits the process when a sequence is not seq_len long, bseq.c:54-57: callers only pass files of equal-length records).
int64_t ref_bseq_read(const char *path, int seq_len, char *rows, int64_t cap)
{
	bseq_file_t *fp = bseq_open(path);
	if (!fp) return -1;
	int n = 0;
	bseq1_t *seqs = bseq_read(fp, &n, seq_len);
	bseq_close(fp);
	for (int i = 0; i < n; ++i) {
		if (i < cap) memcpy(rows + (size_t)i * seq_len, seqs[i].seq, seq_len);
		free(seqs[i].seq);
	}
	free(seqs);
	return n;
}

}
