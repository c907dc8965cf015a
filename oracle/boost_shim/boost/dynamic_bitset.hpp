// TEST INFRASTRUCTURE — not part of the product path.
//
// Header-only stand-in for <boost/dynamic_bitset.hpp>, providing exactly the
// surface the reference sources under /root/reference use:
// ctor(nbits, value), size, operator[] (proxy + const), count, reset,
// set(pos,len,val), <<=, >>=, <<, >>, ~, ^, &, |, |=, &=, ^=, to_ulong.
// Bit semantics follow Boost: bit i lives in block i/64; shifts move bits
// towards higher (<<) / lower (>>) indices and drop what leaves [0,size);
// unused high bits of the last block are always kept zero.
//
// This container has no Boost installation; the unmodified reference is
// compiled against this file only to build the oracle binaries in
// oracle/_ref/ (see oracle/Makefile).
#ifndef RB_SHIM_DYNAMIC_BITSET_HPP
#define RB_SHIM_DYNAMIC_BITSET_HPP

#include <algorithm>
#include <cstddef>
#include <cstdint>
#include <stdexcept>
#include <vector>

namespace boost {

template <typename Block = unsigned long, typename Alloc = std::allocator<Block>>
class dynamic_bitset {
    static_assert(sizeof(Block) == 8, "shim assumes 64-bit blocks");

public:
    typedef std::size_t size_type;
    static const size_type npos = static_cast<size_type>(-1);

    class reference {
        Block &blk_;
        Block mask_;

    public:
        reference(Block &b, unsigned bit) : blk_(b), mask_(Block(1) << bit) {}
        operator bool() const { return (blk_ & mask_) != 0; }
        bool operator~() const { return (blk_ & mask_) == 0; }
        reference &operator=(bool v) {
            if (v) blk_ |= mask_;
            else blk_ &= ~mask_;
            return *this;
        }
        reference &operator=(const reference &o) { return *this = static_cast<bool>(o); }
        reference &operator|=(bool v) { if (v) blk_ |= mask_; return *this; }
        reference &operator&=(bool v) { if (!v) blk_ &= ~mask_; return *this; }
        reference &operator^=(bool v) { if (v) blk_ ^= mask_; return *this; }
        reference &flip() { blk_ ^= mask_; return *this; }
    };

    dynamic_bitset() : nbits_(0) {}
    explicit dynamic_bitset(size_type nbits, unsigned long long value = 0)
        : nbits_(nbits), w_((nbits + 63) / 64, Block(0)) {
        if (!w_.empty()) { w_[0] = static_cast<Block>(value); trim(); }
    }

    size_type size() const { return nbits_; }
    size_type num_blocks() const { return w_.size(); }
    bool empty() const { return nbits_ == 0; }

    reference operator[](size_type pos) { return reference(w_[pos >> 6], unsigned(pos & 63)); }
    bool operator[](size_type pos) const { return (w_[pos >> 6] >> (pos & 63)) & 1; }
    bool test(size_type pos) const { return (*this)[pos]; }

    size_type count() const {
        size_type c = 0;
        for (Block b : w_) c += __builtin_popcountll(b);
        return c;
    }
    bool any() const { for (Block b : w_) if (b) return true; return false; }
    bool none() const { return !any(); }

    dynamic_bitset &reset() { std::fill(w_.begin(), w_.end(), Block(0)); return *this; }
    dynamic_bitset &reset(size_type pos) { w_[pos >> 6] &= ~(Block(1) << (pos & 63)); return *this; }
    dynamic_bitset &set() { std::fill(w_.begin(), w_.end(), ~Block(0)); trim(); return *this; }
    dynamic_bitset &set(size_type pos, bool val = true) {
        if (val) w_[pos >> 6] |= Block(1) << (pos & 63);
        else w_[pos >> 6] &= ~(Block(1) << (pos & 63));
        return *this;
    }
    // range set (Boost >= 1.66): bits [pos, pos+len)
    dynamic_bitset &set(size_type pos, size_type len, bool val) {
        for (size_type i = pos; i < pos + len; ++i) set(i, val);
        return *this;
    }
    dynamic_bitset &flip() { for (Block &b : w_) b = ~b; trim(); return *this; }

    void resize(size_type nbits, bool value = false) {
        size_type old = nbits_;
        w_.resize((nbits + 63) / 64, value ? ~Block(0) : Block(0));
        nbits_ = nbits;
        if (value) for (size_type i = old; i < nbits && (i & 63); ++i) set(i, true);
        trim();
    }
    void clear() { w_.clear(); nbits_ = 0; }
    void push_back(bool bit) { resize(nbits_ + 1); set(nbits_ - 1, bit); }

    dynamic_bitset &operator<<=(size_type n) {
        if (n >= nbits_) return reset();
        if (n == 0) return *this;
        const size_type nb = w_.size(), div = n >> 6, r = n & 63;
        if (r == 0) {
            for (size_type i = nb; i-- > div;) w_[i] = w_[i - div];
        } else {
            for (size_type i = nb; i-- > div + 1;)
                w_[i] = (w_[i - div] << r) | (w_[i - div - 1] >> (64 - r));
            w_[div] = w_[0] << r;
        }
        std::fill(w_.begin(), w_.begin() + div, Block(0));
        trim();
        return *this;
    }
    dynamic_bitset &operator>>=(size_type n) {
        if (n >= nbits_) return reset();
        if (n == 0) return *this;
        const size_type nb = w_.size(), div = n >> 6, r = n & 63;
        if (r == 0) {
            for (size_type i = 0; i + div < nb; ++i) w_[i] = w_[i + div];
        } else {
            for (size_type i = 0; i + div + 1 < nb; ++i)
                w_[i] = (w_[i + div] >> r) | (w_[i + div + 1] << (64 - r));
            w_[nb - div - 1] = w_[nb - 1] >> r;
        }
        std::fill(w_.begin() + (nb - div), w_.end(), Block(0));
        return *this;
    }
    dynamic_bitset operator<<(size_type n) const { dynamic_bitset r(*this); r <<= n; return r; }
    dynamic_bitset operator>>(size_type n) const { dynamic_bitset r(*this); r >>= n; return r; }
    dynamic_bitset operator~() const { dynamic_bitset r(*this); r.flip(); return r; }

    dynamic_bitset &operator&=(const dynamic_bitset &o) { for (size_type i = 0; i < w_.size(); ++i) w_[i] &= o.w_[i]; return *this; }
    dynamic_bitset &operator|=(const dynamic_bitset &o) { for (size_type i = 0; i < w_.size(); ++i) w_[i] |= o.w_[i]; return *this; }
    dynamic_bitset &operator^=(const dynamic_bitset &o) { for (size_type i = 0; i < w_.size(); ++i) w_[i] ^= o.w_[i]; return *this; }

    unsigned long to_ulong() const {
        for (size_type i = 1; i < w_.size(); ++i)
            if (w_[i]) throw std::overflow_error("dynamic_bitset::to_ulong overflow");
        return w_.empty() ? 0ul : static_cast<unsigned long>(w_[0]);
    }

    bool operator==(const dynamic_bitset &o) const { return nbits_ == o.nbits_ && w_ == o.w_; }
    bool operator!=(const dynamic_bitset &o) const { return !(*this == o); }

    // shim-only accessors (used by oracle/ tooling, never by the reference)
    const std::vector<Block> &shim_blocks() const { return w_; }
    std::vector<Block> &shim_blocks() { return w_; }

private:
    void trim() {
        if (nbits_ & 63) w_.back() &= (Block(1) << (nbits_ & 63)) - 1;
    }
    size_type nbits_;
    std::vector<Block> w_;
};

template <typename B, typename A>
inline dynamic_bitset<B, A> operator&(const dynamic_bitset<B, A> &a, const dynamic_bitset<B, A> &b) { dynamic_bitset<B, A> r(a); r &= b; return r; }
template <typename B, typename A>
inline dynamic_bitset<B, A> operator|(const dynamic_bitset<B, A> &a, const dynamic_bitset<B, A> &b) { dynamic_bitset<B, A> r(a); r |= b; return r; }
template <typename B, typename A>
inline dynamic_bitset<B, A> operator^(const dynamic_bitset<B, A> &a, const dynamic_bitset<B, A> &b) { dynamic_bitset<B, A> r(a); r ^= b; return r; }

}  // namespace boost

#endif
