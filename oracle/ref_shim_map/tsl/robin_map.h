// Stand-in for tsl::robin_map (oracle build only; absent from the image): the interface VoxelHashMap uses.
// Iteration is in insertion order and tolerates erase(key) of any element during a range-for: the erased entry is
// only marked dead.  (The real container's iterator behaviour after such an erase is its own business -- the
// reference does erase inside its range-for, VoxelHashMap.cpp:94-100; this stand-in implements what that loop intends.)
#pragma once
#include <cstddef>
#include <unordered_map>
#include <utility>
#include <vector>
namespace tsl {
template <class K, class V, class H>
class robin_map {
  struct Entry {
    std::pair<K, V> kv;
    bool alive;
  };
  std::vector<Entry> e_;
  std::unordered_map<K, size_t, H> idx_;

 public:
  class iterator {
   public:
    robin_map *m = nullptr;
    size_t i = 0;
    void skip() { while (i < m->e_.size() && !m->e_[i].alive) i++; }
    std::pair<K, V> &operator*() const { return m->e_[i].kv; }
    iterator &operator++() { i++; skip(); return *this; }
    bool operator!=(const iterator &o) const { return i != o.i; }
    bool operator==(const iterator &o) const { return i == o.i; }
    V &value() const { return m->e_[i].kv.second; }
  };
  iterator begin() { iterator it{this, 0}; it.skip(); return it; }
  iterator end() { return iterator{this, e_.size()}; }
  iterator find(const K &k) {
    auto f = idx_.find(k);
    return f == idx_.end() ? end() : iterator{this, f->second};
  }
  bool contains(const K &k) const { return idx_.count(k) != 0; }
  void insert(const std::pair<K, V> &kv) {
    if (idx_.count(kv.first)) return;
    idx_[kv.first] = e_.size();
    e_.push_back(Entry{kv, true});
  }
  void erase(const K &k) {
    auto f = idx_.find(k);
    if (f == idx_.end()) return;
    e_[f->second].alive = false;
    idx_.erase(f);
  }
  bool empty() const { return idx_.empty(); }
  size_t size() const { return idx_.size(); }
  void clear() { e_.clear(); idx_.clear(); }
};
}  // namespace tsl
