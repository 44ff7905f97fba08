// tests/stubs/config.h -- TEST INFRASTRUCTURE.  The smallest ConfigManager / json-like node that lets the reference's
// radio-module demodulator wrappers (decoder_modules/radio/src/demodulators/*.h) compile and run unchanged against the
// adapter headers: conf[name][key] chains, contains(), assignment from and conversion to bool / float / string.
#pragma once
#include <map>
#include <string>

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
    operator bool() const { return num != 0.0; }
    operator float() const { return (float)num; }
    operator double() const { return num; }
    operator int() const { return (int)num; }
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
