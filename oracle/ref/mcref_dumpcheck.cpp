// TEST INFRASTRUCTURE (oracle/_ref build only; never linked into the product).
// CPU check of dropin/mcb_dump_writer.h — the marshalling half of N2 (which file gets which bytes) — against the UNMODIFIED
// reference: this file defines kt_dump_for / kt_dump_pe_for on top of the writer, with an encoder that calls the oracle's
// restatement of print_encode (oracle/mc_oracle.c: mco_print_encode) on the host strings.  build_ref.sh links it with all the
// reference's objects, kthread_dump[_pe].o having its own kt_dump_[pe_]for weakened, into minicom_ref_L<L>_<mode>_dumpcheck; the
// test runs that binary and the plain reference on the same reads and compares the directories byte for byte, no GPU involved.
// The product's encoder (dropin/mcb_dropin.cpp) differs only in calling the device kernel (mcb_dump_encode).
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <string>
#include <vector>
#include "breads.h"
#include "kvec.h"
#include "config.h"
#include "mcb_dump_writer.h"

extern "C" int mco_print_encode(const char *read, int dir, const char *ref_window, int L, char *out);

struct OracleEncoder : McbDumpEncoder {
	std::vector<uint64_t> off;
	std::string enc;
	int encode(const uint64_t *members, const uint64_t *moff, const char *refs, const uint64_t *roff, uint64_t nc, const uint64_t **enc_off, const char **enc_out)
	{
		const int L = reads->seq_len;
		off.assign(1, 0); enc.clear();
		std::vector<char> tmp((size_t)L + 1), out((size_t)L + 2);
		for (uint64_t c = 0; c < nc; ++c)
			for (uint64_t k = moff[c]; k < moff[c + 1]; ++k) {
				const uint64_t y = members[k];
				const uint32_t rid = (uint32_t)(y >> 32);
				const int pos = (int)((uint32_t)y >> 1), dir = (int)(y & 1);
				bseq1_t *seq = &reads->seq[rid];
				memcpy(tmp.data(), seq->seq, (size_t)L);
				uint32_v *np = (uint32_v*)seq->n_pos;                       // put N back (kthread_dump.c:70-75)
				if (np) for (size_t i = 0; i < np->n; ++i) tmp[np->a[i]] = 'N';
				const int n = mco_print_encode(tmp.data(), dir, refs + roff[c] + pos, L, out.data());
				enc.append(out.data(), (size_t)n);
				off.push_back(enc.size());
			}
		*enc_off = off.data(); *enc_out = enc.data();
		return 0;
	}
};

#ifdef _PE
void kt_dump_pe_for(int n_threads_, reads_t *r, int index)
#else
void kt_dump_for(int n_threads_, reads_t *r, int index)
#endif
{
	OracleEncoder E;
	mcb_dump_workers(n_threads_, r, index, E);
}
