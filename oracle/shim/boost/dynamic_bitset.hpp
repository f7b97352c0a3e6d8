// Minimal header-only stand-in for boost::dynamic_bitset<> covering exactly what the
// reference uses: ctor(n), operator[] read/write, resize(n, val), size(), count().
// TEST INFRASTRUCTURE ONLY (see string.hpp).
#ifndef HC_ORACLE_SHIM_BOOST_DYNAMIC_BITSET_HPP
#define HC_ORACLE_SHIM_BOOST_DYNAMIC_BITSET_HPP
#include <vector>
#include <cstddef>
#include <memory>
#include <deque>
#include <functional>
#include <algorithm>

namespace boost {
template <class Block = unsigned long>
class dynamic_bitset {
    std::vector<unsigned char> bits_;
public:
    class reference {
        unsigned char& b_;
    public:
        explicit reference(unsigned char& b) : b_(b) {}
        reference& operator=(bool v) { b_ = v ? 1 : 0; return *this; }
        reference& operator=(const reference& o) { b_ = o.b_; return *this; }
        operator bool() const { return b_ != 0; }
        bool operator!() const { return b_ == 0; }
        bool operator~() const { return b_ == 0; }
        reference& flip() { b_ = !b_; return *this; }
    };
    dynamic_bitset() {}
    explicit dynamic_bitset(std::size_t n, unsigned long = 0) : bits_(n, 0) {}
    reference operator[](std::size_t i) { return reference(bits_[i]); }
    bool operator[](std::size_t i) const { return bits_[i] != 0; }
    bool test(std::size_t i) const { return bits_.at(i) != 0; }
    dynamic_bitset& set(std::size_t i, bool v = true) { bits_.at(i) = v ? 1 : 0; return *this; }
    dynamic_bitset& reset() { std::fill(bits_.begin(), bits_.end(), 0); return *this; }
    void resize(std::size_t n, bool v = false) { bits_.resize(n, v ? 1 : 0); }
    std::size_t size() const { return bits_.size(); }
    std::size_t count() const { std::size_t c = 0; for (auto b : bits_) c += b; return c; }
    bool any() const { return count() > 0; }
    bool none() const { return count() == 0; }
    void push_back(bool v) { bits_.push_back(v ? 1 : 0); }
    void clear() { bits_.clear(); }
};
}  // namespace boost
#endif
