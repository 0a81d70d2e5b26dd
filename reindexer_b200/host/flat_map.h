// Host-side label -> internal row index dictionary (the role of BruteforceSearch::dictExternalToInternal_,
// cpp_src/core/index/float_vector/hnswlib/bruteforce.h:56-59).  Open addressing, linear probing, backward-shift erase;
// 12 bytes per slot at load factor <= 0.5, so a 10M-row shard costs ~400 MB of host RAM and builds in well under a second.
#pragma once
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <new>

namespace rxgpu {

class LabelMap {
public:
	static constexpr uint32_t kNotFound = 0xFFFFFFFFu;

	LabelMap() = default;
	LabelMap(const LabelMap& o) { *this = o; }
	LabelMap& operator=(const LabelMap& o) {
		if (this != &o) {
			release();
			cap_ = o.cap_;
			size_ = o.size_;
			if (cap_) {
				alloc(cap_);
				std::memcpy(keys_, o.keys_, cap_ * sizeof(uint64_t));
				std::memcpy(vals_, o.vals_, cap_ * sizeof(uint32_t));
			}
		}
		return *this;
	}
	~LabelMap() { release(); }

	size_t size() const noexcept { return size_; }
	void reserve(size_t n) {
		size_t want = 16;
		while (want < n * 2) {
			want <<= 1;
		}
		if (want > cap_) {
			rehash(want);
		}
	}
	uint32_t find(uint64_t key) const noexcept {
		if (!cap_) {
			return kNotFound;
		}
		for (size_t s = slot(key);; s = (s + 1) & (cap_ - 1)) {
			if (vals_[s] == kNotFound) {
				return kNotFound;
			}
			if (keys_[s] == key) {
				return vals_[s];
			}
		}
	}
	void put(uint64_t key, uint32_t val) {
		if ((size_ + 1) * 2 > cap_) {
			rehash(cap_ ? cap_ * 2 : 16);
		}
		size_t s = slot(key);
		while (vals_[s] != kNotFound && keys_[s] != key) {
			s = (s + 1) & (cap_ - 1);
		}
		if (vals_[s] == kNotFound) {
			keys_[s] = key;
			++size_;
		}
		vals_[s] = val;
	}
	void erase(uint64_t key) noexcept {
		if (!cap_) {
			return;
		}
		size_t hole = slot(key);
		for (;; hole = (hole + 1) & (cap_ - 1)) {
			if (vals_[hole] == kNotFound) {
				return;
			}
			if (keys_[hole] == key) {
				break;
			}
		}
		vals_[hole] = kNotFound;
		--size_;
		for (size_t s = (hole + 1) & (cap_ - 1); vals_[s] != kNotFound; s = (s + 1) & (cap_ - 1)) {
			const size_t home = slot(keys_[s]);
			const bool between = hole <= s ? (home > hole && home <= s) : (home > hole || home <= s);
			if (!between) {
				keys_[hole] = keys_[s];
				vals_[hole] = vals_[s];
				vals_[s] = kNotFound;
				hole = s;
			}
		}
	}
	size_t allocated_bytes() const noexcept { return cap_ * (sizeof(uint64_t) + sizeof(uint32_t)); }

private:
	static uint64_t mix(uint64_t z) noexcept {
		z += 0x9E3779B97F4A7C15ull;
		z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
		z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
		return z ^ (z >> 31);
	}
	size_t slot(uint64_t key) const noexcept { return size_t(mix(key)) & (cap_ - 1); }
	void alloc(size_t cap) {
		keys_ = static_cast<uint64_t*>(std::malloc(cap * sizeof(uint64_t)));
		vals_ = static_cast<uint32_t*>(std::malloc(cap * sizeof(uint32_t)));
		if (!keys_ || !vals_) {
			throw std::bad_alloc();
		}
	}
	void release() noexcept {
		std::free(keys_);
		std::free(vals_);
		keys_ = nullptr;
		vals_ = nullptr;
		cap_ = size_ = 0;
	}
	void rehash(size_t ncap) {
		uint64_t* ok = keys_;
		uint32_t* ov = vals_;
		const size_t ocap = cap_;
		alloc(ncap);
		std::memset(vals_, 0xFF, ncap * sizeof(uint32_t));
		cap_ = ncap;
		size_ = 0;
		for (size_t s = 0; s < ocap; ++s) {
			if (ov[s] != kNotFound) {
				put(ok[s], ov[s]);
			}
		}
		std::free(ok);
		std::free(ov);
	}

	uint64_t* keys_ = nullptr;
	uint32_t* vals_ = nullptr;
	size_t cap_ = 0;
	size_t size_ = 0;
};

}  // namespace rxgpu
