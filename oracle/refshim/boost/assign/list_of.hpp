// stand-in for <boost/assign/list_of.hpp>: map_list_of(k, v)(k, v)... convertible to std::map
#pragma once
#include <map>
#include <utility>
#include <vector>
namespace boost { namespace assign {
template <class K, class V>
class map_list {
 public:
  map_list &operator()(const K &k, const V &v) { items_.push_back(std::make_pair(k, v)); return *this; }
  template <class MK, class MV, class C, class A>
  operator std::map<MK, MV, C, A>() const {
    std::map<MK, MV, C, A> m;
    for (size_t i = 0; i < items_.size(); i++) m.insert(std::make_pair(MK(items_[i].first), MV(items_[i].second)));
    return m;
  }
 private:
  std::vector<std::pair<K, V> > items_;
};
template <class K, class V>
inline map_list<K, V> map_list_of(const K &k, const V &v) { map_list<K, V> l; l(k, v); return l; }
template <class K>
inline map_list<K, const char *> map_list_of(const K &k, const char *v) { map_list<K, const char *> l; l(k, v); return l; }
} }  // namespace boost::assign
