// mcb_dump_writer.h — N2: the per-contig half of the dump stage, kt_dump_for / kt_dump_pe_for
// (kthread_dump.c:239-362, kthread_dump_pe.c:122-210) with their print_encode / print_pe_encode (kthread_dump.c:33-236,
// kthread_dump_pe.c:35-120), as marshalling around ONE batched call that encodes every member of every contig.
//
// For every thread slot `tid` of reads->clusters[index] the reference opens ref.bin.<tid>, beg_pos.bin.<tid>, dir.bin.<tid>,
// dif_char.txt.<tid> (+ ids.bin.<tid> under ORDER, ids.txt.<tid> under _PE) and, contig by contig: sorts the members
// (cmpcluster2; cmpcluster3 under ORDER / _PE), appends the consensus 2 bits per base (DNA_push, carried across contigs),
// the member count (u32) and the position deltas (u16), one direction bit per member (bit_push, carried across contigs), one
// line of dif_char.txt per member (the diff encoding: what McbDumpEncoder produces), and the read ids (ORDER: delta-coded
// against the previous member when the position repeats; _PE: "<file> <rid>" text lines).
// The product passes an encoder that calls mcb_dump_encode (the device kernel); a CPU harness under oracle/ref/ passes the
// oracle's restatement to check this file's byte layout against the unmodified reference without a GPU.
#pragma once
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <stdint.h>
#include <string>
#include <vector>

struct McbDumpEncoder {
	// members contig-major (y = rid<<32 | pos<<1 | dir), moff[nc+1], refs concatenated, roff[nc+1] -> enc_off[n+1], enc
	virtual int encode(const uint64_t *members, const uint64_t *moff, const char *refs, const uint64_t *roff, uint64_t nc,
	                   const uint64_t **enc_off, const char **enc) = 0;
	virtual ~McbDumpEncoder() {}
};

static void mcb_dump_fail(const char *what, const char *name)
{
	fprintf(stderr, "minicom_b200: dump stage: %s %s\n", what, name);
	exit(1);
}

// everything the per-thread worker of the reference writes for slot `tid`
static void mcb_dump_slot(reads_t *r, int index, int tid, McbDumpEncoder &E)
{
	cluster_v *cv = &r->clusters[index][tid];
	const size_t nc = cv->n;
	std::vector<uint64_t> moff(nc + 1, 0), roff(nc + 1, 0);
	for (size_t c = 0; c < nc; ++c) {
		cluster_t *p = &cv->a[c];
#if defined(ORDER) || defined(_PE)
		qsort(p->a, p->n, sizeof(uint64_t), cmpcluster3);            // kthread_dump.c:34, kthread_dump_pe.c:36
#else
		qsort(p->a, p->n, sizeof(uint64_t), cmpcluster2);            // kthread_dump.c:127
#endif
		moff[c + 1] = moff[c] + p->n;
		roff[c + 1] = roff[c] + strlen(p->ref);
	}
	std::vector<uint64_t> members(moff[nc]);
	std::string refs;
	refs.reserve(roff[nc]);
	for (size_t c = 0; c < nc; ++c) {
		cluster_t *p = &cv->a[c];
		if (p->n) memcpy(&members[moff[c]], p->a, p->n * sizeof(uint64_t));
		refs.append(p->ref, roff[c + 1] - roff[c]);
	}
	const uint64_t *eoff = 0;
	const char *enc = 0;
	if (nc && E.encode(members.data(), moff.data(), refs.data(), roff.data(), nc, &eoff, &enc)) mcb_dump_fail("encoding failed for thread slot", std::to_string(tid).c_str());

	char name[4200];
	auto open_file = [&](const char *stem) {
		snprintf(name, sizeof name, "%s/%s.%d", folder, stem, tid);
		FILE *f = fopen(name, "wb");
		if (!f) mcb_dump_fail("cannot create", name);
		return f;
	};
	FILE *fpref = open_file("ref.bin"), *fppos = open_file("beg_pos.bin"), *fpdir = open_file("dir.bin"), *fpdif = open_file("dif_char.txt");
#if defined(_PE)
	FILE *fpids = open_file("ids.txt");
#elif defined(ORDER)
	FILE *fpids = open_file("ids.bin");
#endif
	// ---- consensus strings, 2 bits per base, four bases per byte, carried across the contigs of the slot (DNA_push, breads.h:232)
	{
		std::vector<uint8_t> out;
		out.reserve(roff[nc] / 4 + 1);
		unsigned acc = 0, cnt = 0;
		for (uint64_t i = 0; i < roff[nc]; ++i) {
			acc += (unsigned)seq_nt4_table[(uint8_t)refs[i]] << (2 * cnt);
			if (++cnt == 4) { out.push_back((uint8_t)acc); acc = cnt = 0; }
		}
		if (cnt > 0) out.push_back((uint8_t)acc);                     // the worker's final flush (kthread_dump.c:305-307)
		if (!out.empty() && fwrite(out.data(), 1, out.size(), fpref) != out.size()) mcb_dump_fail("cannot write", "ref.bin");
	}
	// ---- per contig: member count, position deltas; per member: direction bit, diff line, ids
	{
		std::vector<uint8_t> pos_out, dir_out, ids_out;
		std::string ids_txt;
		pos_out.reserve(nc * 4 + members.size() * 2);
		dir_out.reserve(members.size() / 8 + 1);
		unsigned dacc = 0, dcnt = 0;
		for (size_t c = 0; c < nc; ++c) {
			const uint32_t num = (uint32_t)(moff[c + 1] - moff[c]);
			const uint8_t *nb = (const uint8_t*)&num;
			pos_out.insert(pos_out.end(), nb, nb + 4);
			int pre_pos = 0;
			uint32_t pre_rid = 0;
			for (uint64_t k = moff[c]; k < moff[c + 1]; ++k) {
				const uint64_t y = members[k];
				const uint32_t rid = (uint32_t)(y >> 32);
				const int pos = (int)((uint32_t)y >> 1), dir = (int)(y & 1);
				const uint16_t posbin = (uint16_t)(pos - pre_pos);
				const uint8_t *pb = (const uint8_t*)&posbin;
				pos_out.insert(pos_out.end(), pb, pb + 2);
#if defined(_PE)
				{ char line[32]; const int n = snprintf(line, sizeof line, "%d %u\n", rid < (uint32_t)half_val ? 0 : 1, rid); ids_txt.append(line, (size_t)n); }   // kthread_dump_pe.c:66-70
#elif defined(ORDER)
				{   // the first read of a contig, or a new begin position: the id itself; same position: the difference (kthread_dump.c:95-105)
					const uint32_t v = (k == moff[c] || posbin > 0) ? rid : rid - pre_rid;
					const uint8_t *vb = (const uint8_t*)&v;
					ids_out.insert(ids_out.end(), vb, vb + 4);
				}
#endif
				dacc += (unsigned)dir << dcnt;                          // bit_push (breads.h:241)
				if (++dcnt == 8) { dir_out.push_back((uint8_t)dacc); dacc = dcnt = 0; }
				pre_pos = pos; pre_rid = rid;
			}
		}
		(void)ids_out; (void)ids_txt;
		if (dcnt > 0) dir_out.push_back((uint8_t)dacc);
		if (!pos_out.empty() && fwrite(pos_out.data(), 1, pos_out.size(), fppos) != pos_out.size()) mcb_dump_fail("cannot write", "beg_pos.bin");
		if (!dir_out.empty() && fwrite(dir_out.data(), 1, dir_out.size(), fpdir) != dir_out.size()) mcb_dump_fail("cannot write", "dir.bin");
#if defined(_PE)
		if (!ids_txt.empty() && fwrite(ids_txt.data(), 1, ids_txt.size(), fpids) != ids_txt.size()) mcb_dump_fail("cannot write", "ids.txt");
#elif defined(ORDER)
		if (!ids_out.empty() && fwrite(ids_out.data(), 1, ids_out.size(), fpids) != ids_out.size()) mcb_dump_fail("cannot write", "ids.bin");
#endif
	}
	// ---- dif_char.txt: one line per member
	if (members.size()) {
		std::string lines;
		lines.reserve((size_t)eoff[members.size()] + members.size());
		for (uint64_t k = 0; k < members.size(); ++k) { lines.append(enc + eoff[k], (size_t)(eoff[k + 1] - eoff[k])); lines.push_back('\n'); }
		if (fwrite(lines.data(), 1, lines.size(), fpdif) != lines.size()) mcb_dump_fail("cannot write", "dif_char.txt");
	}
	fclose(fpref); fclose(fppos); fclose(fpdir); fclose(fpdif);
#if defined(ORDER) || defined(_PE)
	fclose(fpids);
#endif
}

static void mcb_dump_workers(int n_threads_, reads_t *r, int index, McbDumpEncoder &E)
{
	for (int tid = 0; tid < n_threads_; ++tid) mcb_dump_slot(r, index, tid, E);   // one file set per slot, like one worker thread each
}
