// stand-in for <boost/bimap.hpp>: two std::maps kept in step; the subset of the interface the reference uses
#pragma once
#include <map>
#include <stdexcept>
namespace boost {
namespace bimaps { namespace tags {
template <class T, class Tag> struct tagged { typedef T value_type; typedef Tag tag; };
} }
namespace bimap_detail {
template <class T> struct untag { typedef T type; };
template <class T, class Tag> struct untag<bimaps::tags::tagged<T, Tag> > { typedef T type; };
}
template <class L, class R>
class bimap {
 public:
  typedef typename bimap_detail::untag<L>::type left_key;
  typedef typename bimap_detail::untag<R>::type right_key;
  typedef std::map<left_key, right_key> left_map;
  typedef std::map<right_key, left_key> right_map;
  typedef typename left_map::const_iterator left_const_iterator;
  typedef typename right_map::const_iterator right_const_iterator;
  struct value_type {
    left_key left; right_key right;
    value_type(const left_key &l, const right_key &r) : left(l), right(r) {}
  };
  typedef value_type relation;
  left_map left;
  right_map right;
  void insert(const value_type &v) { left.insert(std::make_pair(v.left, v.right)); right.insert(std::make_pair(v.right, v.left)); }
  template <class Tag> const right_map &by() const { return right; }
};
}  // namespace boost
