// TEST INFRASTRUCTURE — minimal stand-in for <boost/property_tree/ptree.hpp>: the surface
// Solver::read_parameterfile uses (src/solver.cpp:155-299): get<T>(key), get<T>(key, default), get_child(key), size(),
// range-for over children yielding pair<string, ptree> with .second.data().  Every JSON value is kept as its text, as
// boost::property_tree does; get<T> converts through an istream like boost's stream_translator (bool: true / false / 0 / 1).
#pragma once
#include <sstream>
#include <stdexcept>
#include <string>
#include <utility>
#include <vector>

namespace boost { namespace property_tree {

class ptree_error : public std::runtime_error { public: explicit ptree_error(const std::string & w) : std::runtime_error(w) {} };
class ptree_bad_path : public ptree_error { public: explicit ptree_bad_path(const std::string & w) : ptree_error(w) {} };
class ptree_bad_data : public ptree_error { public: explicit ptree_bad_data(const std::string & w) : ptree_error(w) {} };

class ptree {
public:
    typedef std::pair<std::string, ptree> value_type;
    typedef std::vector<value_type>::iterator iterator;
    typedef std::vector<value_type>::const_iterator const_iterator;
private:
    std::string m_data;
    std::vector<value_type> m_children;

    template <class T> static bool convert(const std::string & s, T & out)
    {
        std::istringstream iss(s);
        iss >> out;
        if (iss.fail()) return false;
        iss >> std::ws;
        return iss.eof();
    }
    static bool convert(const std::string & s, std::string & out) { out = s; return true; }
    static bool convert(const std::string & s, bool & out)
    {
        if (s == "true" || s == "1") { out = true; return true; }
        if (s == "false" || s == "0") { out = false; return true; }
        return false;
    }
    const ptree * find(const std::string & key) const
    {
        for (const auto & c : m_children) if (c.first == key) return &c.second;
        return nullptr;
    }
public:
    ptree() {}
    explicit ptree(const std::string & d) : m_data(d) {}
    const std::string & data() const { return m_data; }
    std::string & data() { return m_data; }
    size_t size() const { return m_children.size(); }
    iterator begin() { return m_children.begin(); }
    iterator end() { return m_children.end(); }
    const_iterator begin() const { return m_children.begin(); }
    const_iterator end() const { return m_children.end(); }
    void push_back(const value_type & v) { m_children.push_back(v); }

    ptree & get_child(const std::string & key)
    {
        for (auto & c : m_children) if (c.first == key) return c.second;
        throw ptree_bad_path("No such node (" + key + ")");
    }
    template <class T> T get(const std::string & key) const
    {
        const ptree * c = find(key);
        if (!c) throw ptree_bad_path("No such node (" + key + ")");
        T v;
        if (!convert(c->m_data, v)) throw ptree_bad_data("conversion of data to type \"" + std::string(typeid(T).name()) + "\" failed");
        return v;
    }
    template <class T> T get(const std::string & key, const T & def) const
    {
        const ptree * c = find(key);
        if (!c) return def;
        T v;
        return convert(c->m_data, v) ? v : def;
    }
    // get<std::string>("SPHType", "ssph"): the default is a string literal (src/solver.cpp:205,248)
    template <class T> T get(const std::string & key, const char * def) const { return get<T>(key, T(def)); }
};

} }
