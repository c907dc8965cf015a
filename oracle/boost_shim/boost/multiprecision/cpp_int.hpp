// TEST INFRASTRUCTURE — not part of the product path.
//
// Header-only stand-in for <boost/multiprecision/cpp_int.hpp>: only
// boost::multiprecision::uint256_t with the operations the reference uses
// (SURVEY.md §8c): zero default ctor, construction from built-in integers,
// <<=, >>=, <<, >>, &, |, ^, &=, |=, ==, !=, <, and std::hash (looked up,
// never iterated, so the hash function does not influence results).
#ifndef RB_SHIM_CPP_INT_HPP
#define RB_SHIM_CPP_INT_HPP

#include <cstddef>
#include <cstdint>
#include <functional>
#include <ostream>
#include <type_traits>

namespace boost {
namespace multiprecision {

struct uint256_t {
    uint64_t w[4];

    uint256_t() : w{0, 0, 0, 0} {}
    template <typename T, typename = typename std::enable_if<std::is_integral<T>::value>::type>
    uint256_t(T v) : w{static_cast<uint64_t>(v), 0, 0, 0} {
        if (std::is_signed<T>::value && v < 0) w[1] = w[2] = w[3] = ~uint64_t(0);
    }

    uint256_t &operator<<=(unsigned n) {
        if (n >= 256) { w[0] = w[1] = w[2] = w[3] = 0; return *this; }
        unsigned div = n >> 6, r = n & 63;
        uint64_t o[4] = {0, 0, 0, 0};
        for (int i = 3; i >= int(div); --i) {
            o[i] = w[i - div] << r;
            if (r && i - int(div) - 1 >= 0) o[i] |= w[i - div - 1] >> (64 - r);
        }
        for (int i = 0; i < 4; ++i) w[i] = o[i];
        return *this;
    }
    uint256_t &operator>>=(unsigned n) {
        if (n >= 256) { w[0] = w[1] = w[2] = w[3] = 0; return *this; }
        unsigned div = n >> 6, r = n & 63;
        uint64_t o[4] = {0, 0, 0, 0};
        for (unsigned i = 0; i + div < 4; ++i) {
            o[i] = w[i + div] >> r;
            if (r && i + div + 1 < 4) o[i] |= w[i + div + 1] << (64 - r);
        }
        for (int i = 0; i < 4; ++i) w[i] = o[i];
        return *this;
    }
    uint256_t &operator&=(const uint256_t &o) { for (int i = 0; i < 4; ++i) w[i] &= o.w[i]; return *this; }
    uint256_t &operator|=(const uint256_t &o) { for (int i = 0; i < 4; ++i) w[i] |= o.w[i]; return *this; }
    uint256_t &operator^=(const uint256_t &o) { for (int i = 0; i < 4; ++i) w[i] ^= o.w[i]; return *this; }

    explicit operator bool() const { return (w[0] | w[1] | w[2] | w[3]) != 0; }
    explicit operator uint64_t() const { return w[0]; }
    explicit operator int() const { return static_cast<int>(w[0]); }
    explicit operator unsigned() const { return static_cast<unsigned>(w[0]); }
};

template <typename I, typename = typename std::enable_if<std::is_integral<I>::value>::type>
inline uint256_t operator<<(uint256_t a, I n) { a <<= static_cast<unsigned>(n); return a; }
template <typename I, typename = typename std::enable_if<std::is_integral<I>::value>::type>
inline uint256_t operator>>(uint256_t a, I n) { a >>= static_cast<unsigned>(n); return a; }
inline uint256_t operator&(uint256_t a, const uint256_t &b) { a &= b; return a; }
inline uint256_t operator|(uint256_t a, const uint256_t &b) { a |= b; return a; }
inline uint256_t operator^(uint256_t a, const uint256_t &b) { a ^= b; return a; }
inline bool operator==(const uint256_t &a, const uint256_t &b) {
    return a.w[0] == b.w[0] && a.w[1] == b.w[1] && a.w[2] == b.w[2] && a.w[3] == b.w[3];
}
inline bool operator!=(const uint256_t &a, const uint256_t &b) { return !(a == b); }
inline bool operator<(const uint256_t &a, const uint256_t &b) {
    for (int i = 3; i >= 0; --i) if (a.w[i] != b.w[i]) return a.w[i] < b.w[i];
    return false;
}
inline bool operator>(const uint256_t &a, const uint256_t &b) { return b < a; }
inline bool operator<=(const uint256_t &a, const uint256_t &b) { return !(b < a); }
inline bool operator>=(const uint256_t &a, const uint256_t &b) { return !(a < b); }
inline std::ostream &operator<<(std::ostream &os, const uint256_t &v) {
    char buf[80];
    snprintf(buf, sizeof buf, "0x%016llx%016llx%016llx%016llx", (unsigned long long)v.w[3],
             (unsigned long long)v.w[2], (unsigned long long)v.w[1], (unsigned long long)v.w[0]);
    return os << buf;
}

typedef uint256_t cpp_int;

}  // namespace multiprecision
}  // namespace boost

namespace std {
template <>
struct hash<boost::multiprecision::uint256_t> {
    size_t operator()(const boost::multiprecision::uint256_t &v) const noexcept {
        uint64_t h = 0x9e3779b97f4a7c15ull;
        for (int i = 0; i < 4; ++i) { h ^= v.w[i] + 0x9e3779b97f4a7c15ull + (h << 6) + (h >> 2); }
        return static_cast<size_t>(h);
    }
};
}  // namespace std

#endif
