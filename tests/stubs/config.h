// tests/stubs/config.h -- TEST INFRASTRUCTURE.  The smallest ConfigManager / json-like node that lets the reference's
// radio-module demodulator wrappers (decoder_modules/radio/src/demodulators/*.h) compile and run unchanged against the
// adapter headers: conf[name][key] chains, contains(), assignment from and conversion to bool / float / string.
#pragma once
#include <map>
#include <string>
#include <type_traits>

class ConfNode {
public:
    ConfNode& operator[](const std::string& k) { return kids[k]; }
    ConfNode& operator[](const char* k) { return kids[k]; }
    bool contains(const std::string& k) const { return kids.count(k) != 0; }
    ConfNode& operator=(bool v) { num = v ? 1.0 : 0.0; return *this; }
    ConfNode& operator=(float v) { num = v; return *this; }
    ConfNode& operator=(double v) { num = v; return *this; }
    ConfNode& operator=(int v) { num = v; return *this; }
    ConfNode& operator=(const std::string& v) { str = v; return *this; }
    // numbers convert to any arithmetic type but the character types (so that `std::string s = node` has one reading)
    template <class T, class = typename std::enable_if<std::is_arithmetic<T>::value && !std::is_same<T, char>::value &&
                                                         !std::is_same<T, signed char>::value && !std::is_same<T, unsigned char>::value>::type>
    operator T() const { return (T)(std::is_same<T, bool>::value ? (num != 0.0) : num); }
    operator std::string() const { return str; }
private:
    std::map<std::string, ConfNode> kids;
    double num = 0.0;
    std::string str;
};

class ConfigManager {
public:
    void acquire() {}
    void release(bool modified = false) { (void)modified; }
    ConfNode conf;
};
