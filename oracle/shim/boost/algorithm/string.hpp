// Minimal header-only stand-in for the few Boost.StringAlgo entry points the
// reference sources use (trim_if, is_any_of, split + token_compress_on,
// to_upper_copy).  TEST INFRASTRUCTURE ONLY: it exists so that the unmodified
// reference .cpp files under /root/reference/src compile in an image that has
// no Boost headers.  Not part of the product, not derived from Boost sources.
#ifndef HC_ORACLE_SHIM_BOOST_STRING_HPP
#define HC_ORACLE_SHIM_BOOST_STRING_HPP
#include <string>
#include <vector>
#include <cctype>
#include <memory>
#include <deque>
#include <functional>
#include <algorithm>

namespace boost {

struct is_any_of {
    std::string set;
    is_any_of(const char* s) : set(s) {}
    is_any_of(const std::string& s) : set(s) {}
    bool operator()(char c) const { return set.find(c) != std::string::npos; }
};

enum token_compress_mode_type { token_compress_on, token_compress_off };

template <class Pred>
inline void trim_if(std::string& s, Pred p) {
    size_t b = 0, e = s.size();
    while (b < e && p(s[b])) ++b;
    while (e > b && p(s[e - 1])) --e;
    s = s.substr(b, e - b);
}

inline std::string to_upper_copy(const std::string& s) {
    std::string r(s);
    for (size_t i = 0; i < r.size(); ++i) r[i] = (char)std::toupper((unsigned char)r[i]);
    return r;
}

namespace algorithm {
template <class Pred>
inline std::vector<std::string>& split(std::vector<std::string>& out, const std::string& in, Pred p,
                                       token_compress_mode_type mode = token_compress_off) {
    out.clear();
    std::string cur;
    size_t i = 0, n = in.size();
    while (true) {
        cur.clear();
        while (i < n && !p(in[i])) cur.push_back(in[i++]);
        out.push_back(cur);
        if (i >= n) break;
        ++i;  // skip one separator
        if (mode == token_compress_on) while (i < n && p(in[i])) ++i;
        if (i >= n) { out.push_back(std::string()); break; }
    }
    return out;
}
}  // namespace algorithm
using algorithm::split;

}  // namespace boost
#endif
