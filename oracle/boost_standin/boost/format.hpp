// TEST INFRASTRUCTURE — minimal stand-in for <boost/format.hpp>: `(boost::format("/%05d.dat") % n).str()` with
// integer `%d` / `%0Nd` fields, which is all mitchiinaga/sphcode uses (src/output.cpp:55, src/logger.cpp:43).
#pragma once
#include <cstdio>
#include <string>
#include <vector>

namespace boost {
class format {
    std::string fmt;
    std::vector<long long> args;
public:
    explicit format(const char * f) : fmt(f) {}
    explicit format(const std::string & f) : fmt(f) {}
    template <class T> format & operator%(const T & v) { args.push_back((long long)v); return *this; }
    std::string str() const
    {
        std::string out;
        size_t a = 0;
        for (size_t i = 0; i < fmt.size(); ++i) {
            if (fmt[i] != '%') { out += fmt[i]; continue; }
            if (i + 1 < fmt.size() && fmt[i + 1] == '%') { out += '%'; ++i; continue; }
            size_t j = i + 1;
            std::string spec = "%";
            while (j < fmt.size() && (fmt[j] == '0' || (fmt[j] >= '1' && fmt[j] <= '9'))) spec += fmt[j++];
            if (j < fmt.size() && fmt[j] == 'd') {
                char buf[64];
                std::snprintf(buf, sizeof(buf), (spec + "lld").c_str(), a < args.size() ? args[a] : 0LL);
                ++a;
                out += buf;
                i = j;
            } else out += fmt[i];
        }
        return out;
    }
};
}
