// TEST INFRASTRUCTURE — minimal stand-in for <boost/any.hpp>, only what mitchiinaga/sphcode uses
// (include/solver.hpp:52 `std::unordered_map<std::string, boost::any>`, src/sample/*.cpp `boost::any_cast<int>`).
// Boost is not installed in this image and there is no network; the arithmetic of the hot path does not live in Boost.
#pragma once
#include <memory>
#include <typeinfo>
#include <stdexcept>

namespace boost {
class bad_any_cast : public std::bad_cast {
public:
    const char * what() const noexcept override { return "boost::bad_any_cast: failed conversion using boost::any_cast"; }
};
class any {
    struct base { virtual ~base() {} virtual const std::type_info & type() const = 0; };
    template <class T> struct holder : base {
        T v;
        explicit holder(const T & x) : v(x) {}
        const std::type_info & type() const override { return typeid(T); }
    };
    std::shared_ptr<base> p;
    template <class T> friend T any_cast(const any &);
public:
    any() {}
    template <class T> any(const T & v) : p(std::make_shared<holder<T>>(v)) {}
    bool empty() const { return !p; }
    const std::type_info & type() const { return p ? p->type() : typeid(void); }
};
template <class T> T any_cast(const any & a)
{
    if (a.type() != typeid(T)) throw bad_any_cast();
    return static_cast<any::holder<T> *>(a.p.get())->v;
}
}
