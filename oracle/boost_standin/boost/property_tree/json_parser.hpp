// TEST INFRASTRUCTURE — minimal stand-in for <boost/property_tree/json_parser.hpp>: read_json(path, ptree &) for
// objects, arrays (children with empty keys), strings, numbers, true / false / null — enough for the parameter files
// of mitchiinaga/sphcode (sample/*/*.json).  Values are stored as their text (numbers verbatim).
#pragma once
#include <cctype>
#include <fstream>
#include <sstream>
#include "ptree.hpp"

namespace boost { namespace property_tree {

class json_parser_error : public ptree_error { public: explicit json_parser_error(const std::string & w) : ptree_error(w) {} };

namespace standin_detail {
struct Parser {
    const std::string & s; size_t i; std::string file;
    void ws() { while (i < s.size() && std::isspace((unsigned char)s[i])) ++i; }
    [[noreturn]] void fail(const char * what) { throw json_parser_error(file + ": " + what + " at offset " + std::to_string(i)); }
    std::string str()
    {
        std::string o;
        ++i;
        while (i < s.size() && s[i] != '"') {
            if (s[i] == '\\' && i + 1 < s.size()) {
                ++i;
                switch (s[i]) { case 'n': o += '\n'; break; case 't': o += '\t'; break; case 'r': o += '\r'; break;
                                case 'b': o += '\b'; break; case 'f': o += '\f'; break; default: o += s[i]; }
            } else o += s[i];
            ++i;
        }
        if (i >= s.size()) fail("unterminated string");
        ++i;
        return o;
    }
    ptree value()
    {
        ws();
        if (i >= s.size()) fail("unexpected end");
        if (s[i] == '{') {
            ptree t;
            ++i; ws();
            if (i < s.size() && s[i] == '}') { ++i; return t; }
            for (;;) {
                ws();
                if (i >= s.size() || s[i] != '"') fail("expected key");
                const std::string k = str();
                ws();
                if (i >= s.size() || s[i] != ':') fail("expected ':'");
                ++i;
                t.push_back(std::make_pair(k, value()));
                ws();
                if (i < s.size() && s[i] == ',') { ++i; continue; }
                if (i < s.size() && s[i] == '}') { ++i; return t; }
                fail("expected ',' or '}'");
            }
        }
        if (s[i] == '[') {
            ptree t;
            ++i; ws();
            if (i < s.size() && s[i] == ']') { ++i; return t; }
            for (;;) {
                t.push_back(std::make_pair(std::string(), value()));
                ws();
                if (i < s.size() && s[i] == ',') { ++i; continue; }
                if (i < s.size() && s[i] == ']') { ++i; return t; }
                fail("expected ',' or ']'");
            }
        }
        if (s[i] == '"') return ptree(str());
        size_t j = i;
        while (j < s.size() && s[j] != ',' && s[j] != '}' && s[j] != ']' && !std::isspace((unsigned char)s[j])) ++j;
        if (j == i) fail("expected value");
        ptree t(s.substr(i, j - i));
        i = j;
        return t;
    }
};
}

inline void read_json(const std::string & filename, ptree & pt)
{
    std::ifstream in(filename);
    if (!in) throw json_parser_error(filename + ": cannot open file");
    std::stringstream ss;
    ss << in.rdbuf();
    const std::string text = ss.str();
    standin_detail::Parser p{text, 0, filename};
    pt = p.value();
}

} }
