// oracle/shim/boost/fusion/fusion_shim.hpp — TEST INFRASTRUCTURE, not product code.
// The handful of Boost.Fusion / Boost.MPL facilities the reference's subdomain.h / domain.h use (heterogeneous vectors of
// boundaries and subdomains, at_c, for_each, push_front, transform over an integer range), on top of std::tuple.
// Every fusion / mpl header the reference includes forwards here.
#ifndef MCB_SHIM_BOOST_FUSION
#define MCB_SHIM_BOOST_FUSION
#include <cstddef>
#include <tuple>
#include "../assert.hpp"      // real Boost headers pull BOOST_ASSERT in transitively; subdomain.h relies on that
#include <type_traits>
#include <utility>

namespace boost {
namespace mpl {
template <class T, T V> struct integral_c {
    static const T value = V;
    typedef integral_c type;
    typedef T value_type;
    operator T() const { return V; }
};
template <class T, T V> const T integral_c<T, V>::value;
template <int V> struct int_ : integral_c<int, V> {};
template <class T, T B, T E> struct range_c {};
}
namespace fusion {

template <class... Ts> struct vector {
    std::tuple<Ts...> t;
    vector() {}
    template <class... Us, class = typename std::enable_if<sizeof...(Us) == sizeof...(Ts) && (sizeof...(Us) > 0)>::type>
    vector(const Us&... us) : t(us...) {}
    vector(const vector& o) : t(o.t) {}
    vector& operator=(const vector& o) { t = o.t; return *this; }
};
#define MCB_FUSION_ALIAS(N) template <class... Ts> using vector##N = vector<Ts...>;
MCB_FUSION_ALIAS(1) MCB_FUSION_ALIAS(2) MCB_FUSION_ALIAS(3) MCB_FUSION_ALIAS(4) MCB_FUSION_ALIAS(5) MCB_FUSION_ALIAS(6)
MCB_FUSION_ALIAS(7) MCB_FUSION_ALIAS(8) MCB_FUSION_ALIAS(9) MCB_FUSION_ALIAS(10) MCB_FUSION_ALIAS(42) MCB_FUSION_ALIAS(50)

namespace result_of {
template <class Seq> struct size;
template <class... Ts> struct size<vector<Ts...> > { typedef mpl::integral_c<int, (int)sizeof...(Ts)> type; };
template <class Seq, int I> struct value_at_c;
template <class... Ts, int I> struct value_at_c<vector<Ts...>, I> {
    typedef typename std::tuple_element<(std::size_t)I, std::tuple<Ts...> >::type type;
};
template <class Seq, class I> struct value_at : value_at_c<Seq, (int)I::value> {};
template <class Seq, int I> struct at_c { typedef typename value_at_c<Seq, I>::type& type; };
template <class Seq, int I> struct at_c<const Seq, I> { typedef const typename value_at_c<Seq, I>::type& type; };
template <class Seq, class T> struct push_front;
template <class... Ts, class T> struct push_front<vector<Ts...>, T> { typedef vector<T, Ts...> type; };
template <class Seq> struct as_vector { typedef Seq type; };
}

template <int I, class... Ts> typename std::tuple_element<(std::size_t)I, std::tuple<Ts...> >::type& at_c(vector<Ts...>& v) {
    return std::get<(std::size_t)I>(v.t);
}
template <int I, class... Ts> const typename std::tuple_element<(std::size_t)I, std::tuple<Ts...> >::type& at_c(const vector<Ts...>& v) {
    return std::get<(std::size_t)I>(v.t);
}

namespace detail {
template <class V, class F, std::size_t... Is> void for_each_impl(V& v, const F& f, std::index_sequence<Is...>) {
    int dummy[] = {0, (f(std::get<Is>(v.t)), 0)...};
    (void)dummy;
}
template <class... Ts, class T, std::size_t... Is>
vector<T, Ts...> push_front_impl(const vector<Ts...>& v, const T& x, std::index_sequence<Is...>) {
    return vector<T, Ts...>(x, std::get<Is>(v.t)...);
}
template <int B, class F, std::size_t... Is>
auto transform_impl(const F& f, std::index_sequence<Is...>) -> vector<decltype(f(mpl::integral_c<int, B + (int)Is>()))...> {
    return vector<decltype(f(mpl::integral_c<int, B + (int)Is>()))...>(f(mpl::integral_c<int, B + (int)Is>())...);
}
}
template <class... Ts, class F> void for_each(vector<Ts...>& v, const F& f) {
    detail::for_each_impl(v, f, std::index_sequence_for<Ts...>());
}
template <class... Ts, class F> void for_each(const vector<Ts...>& v, const F& f) {
    detail::for_each_impl(v, f, std::index_sequence_for<Ts...>());
}
template <class... Ts, class T> vector<T, Ts...> push_front(const vector<Ts...>& v, const T& x) {
    return detail::push_front_impl(v, x, std::index_sequence_for<Ts...>());
}
template <class F, int B, int E>
auto transform(const mpl::range_c<int, B, E>&, const F& f)
    -> decltype(detail::transform_impl<B>(f, std::make_index_sequence<(std::size_t)(E - B)>())) {
    return detail::transform_impl<B>(f, std::make_index_sequence<(std::size_t)(E - B)>());
}
template <class Seq> const Seq& as_vector(const Seq& s) { return s; }

} // namespace fusion
} // namespace boost
#endif
