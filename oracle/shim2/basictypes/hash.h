#pragma once   // TEST INFRASTRUCTURE ONLY: stand-in for src/basictypes/hash.h (state hashes are not exercised by the checkers)
#include <cstdint>
namespace ucoslam {
struct Hash {
    uint64_t v = 0;
    template <typename T> void add(const T&) {}
    template <typename It> void add(It, It) {}
    template <typename T> void operator+=(const T&) {}
    operator uint64_t() const { return v; }
};
}
