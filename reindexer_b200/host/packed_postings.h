// Decoder of the reference's packed posting lists (PackedIdRelVec, the default FtFastConfig Optimization::Memory container) into the
// SoA layout the device merge streams.  Product host code (C++): a restatement of
//   IdRelType::unpackWithoutArrayIdxs      cpp_src/core/ft/idrelset.cc:185-235
//   PackedIdRelVec::iterator (state chain) cpp_src/core/ft/idrelset.h:172-211
//   varint scan / parse                    cpp_src/tools/varint.h:122-175
// The stream is one varint record after another with ids and fields delta-coded against the previous record, and it carries no
// skip pointers: decoding is inherently sequential per list, so it happens once, at upload, on the host (a 5 M-posting list is
// ~30 MB and decodes in tens of milliseconds), not per query.  Records written after PackedIdRelVec saw array data
// (arrayFoundPos_, idrelset.h:226-236) use the layout with array indexes (idrelset.cc:8-62); the device path stores no array
// indexes, so such lists are rejected.
#pragma once
#include <cstdint>
#include <vector>

namespace rxgpu {

struct DecodedPostings {
	std::vector<uint32_t> doc_ids;
	std::vector<uint32_t> pos_begin;  // doc_ids.size() + 1
	std::vector<uint32_t> positions;  // pos | field << 24
};

// returns 0 on success, -1 malformed stream / count mismatch, -2 field or position out of the SoA range (field > 255, pos >= 2^24)
inline int decodePackedPostings(const uint8_t* data, uint64_t len, uint32_t count, DecodedPostings& out) {
	out.doc_ids.clear();
	out.pos_begin.assign(1, 0u);
	out.positions.clear();
	out.doc_ids.reserve(count);
	out.pos_begin.reserve(size_t(count) + 1);
	const uint8_t* p = data;
	const uint8_t* const end = data + len;
	bool ok = true;
	auto get = [&]() -> uint32_t {  // base-128 varint, at most 5 bytes for a uint32
		uint32_t v = 0;
		for (unsigned shift = 0; shift < 35; shift += 7) {
			if (p == end) {
				ok = false;
				return 0;
			}
			const uint8_t b = *p++;
			v |= uint32_t(b & 0x7f) << shift;
			if (!(b & 0x80)) {
				return v;
			}
		}
		ok = false;
		return 0;
	};
	uint32_t lastId = 0, lastField = 0;  // PackedIdRelVec::state
	while (p != end) {
		uint32_t id = get();
		uint32_t head = get();
		const bool idModified = head & 1, fieldIsSame = head & 2, sizeIs1 = head & 4;
		uint32_t shift = head >> 3;
		if (idModified) {
			id += lastId;
		}
		uint32_t field = lastField;
		if (!fieldIsSame) {
			field = get();
		}
		uint32_t size = 1;
		if (!sizeIs1) {
			size = get() + 1;
		}
		if (!ok) {
			return -1;
		}
		auto push = [&](uint32_t pos, uint32_t f) {
			if (f > 0xFFu || pos > 0xFFFFFFu) {
				return false;
			}
			out.positions.push_back(pos | (f << 24));
			return true;
		};
		if (!push(shift, field)) {
			return -2;
		}
		uint32_t pf = field, ps = shift;
		for (uint32_t i = 1; i < size; ++i) {
			uint32_t next = get();
			const bool same = next & 1;
			next >>= 1;
			uint32_t nf = pf;
			if (same) {
				next += ps;
			} else {
				nf = get() + pf;
			}
			if (!ok) {
				return -1;
			}
			if (!push(next, nf)) {
				return -2;
			}
			ps = next;
			pf = nf;
		}
		out.doc_ids.push_back(id);
		out.pos_begin.push_back(uint32_t(out.positions.size()));
		lastId = id;
		lastField = field;  // Pos()[0].field()
	}
	return ok && out.doc_ids.size() == count ? 0 : -1;
}

}  // namespace rxgpu
