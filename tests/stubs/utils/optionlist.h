// tests/stubs/utils/optionlist.h -- TEST INFRASTRUCTURE: the part of OptionList<K, T> (core/src/utils/optionlist.h) the
// radio module's WFM wrapper uses: define, key / value lookups, the zero-separated name list for ImGui::Combo
#pragma once
#include <string>
#include <vector>

template <class K, class T>
class OptionList {
public:
    void define(const K& key, const std::string& name, const T& value) {
        keys.push_back(key); names.push_back(name); values.push_back(value);
        list.clear();
        for (auto& n : names) { list += n; list.push_back('\0'); }
        txt = list.c_str();
    }
    bool keyExists(const K& key) const { return index_of(keys, key) >= 0; }
    int keyId(const K& key) const { return index_of(keys, key); }
    int valueId(const T& value) const { return index_of(values, value); }
    const K& key(int id) const { return keys[(size_t)id]; }
    const T& value(int id) const { return values[(size_t)id]; }
    int size() const { return (int)keys.size(); }
    const char* txt = "";
private:
    template <class V> static int index_of(const std::vector<V>& v, const V& x) {
        for (size_t i = 0; i < v.size(); i++) { if (v[i] == x) { return (int)i; } }
        return -1;
    }
    std::vector<K> keys;
    std::vector<std::string> names;
    std::vector<T> values;
    std::string list;
};
