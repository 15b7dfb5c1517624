#pragma once   // TEST INFRASTRUCTURE ONLY: stand-in for src/basictypes/io_utils.h: raw-POD vector / string streaming (io_utils.h:45-120)
#include <cstdint>
#include <iostream>
#include <string>
#include <vector>
namespace ucoslam {
template <typename T> void toStream__(const std::vector<T>& v, std::ostream& str) {
    uint64_t s = v.size();
    str.write((char*)&s, sizeof(s));
    if (s) str.write((char*)&v[0], sizeof(T) * s);
}
template <typename T> void fromStream__(std::vector<T>& v, std::istream& str) {
    uint64_t s;
    str.read((char*)&s, sizeof(s));
    v.resize(s);
    if (s) str.read((char*)&v[0], sizeof(T) * s);
}
inline void toStream__(const std::string& v, std::ostream& str) {
    uint64_t s = v.size();
    str.write((char*)&s, sizeof(s));
    if (s) str.write(v.data(), s);
}
inline void fromStream__(std::string& v, std::istream& str) {
    uint64_t s;
    str.read((char*)&s, sizeof(s));
    v.resize(s);
    if (s) str.read(&v[0], s);
}
}
