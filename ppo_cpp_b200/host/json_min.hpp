// A small stand-in for the subset of nlohmann::json (vendored by the reference as json.hpp, v3.6.1) that the
// hot path's host classes use: operator[](key) on objects, assignment from numbers / strings / float vectors,
// get<T>(), dump(), parse().  If the real header was included first (INCLUDE_NLOHMANN_JSON_HPP_), this file
// defines nothing — a maintainer dropping these classes into ppo_cpp keeps using the vendored json.hpp.
#ifndef PPO_B200_JSON_MIN_HPP
#define PPO_B200_JSON_MIN_HPP
#ifndef INCLUDE_NLOHMANN_JSON_HPP_

#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <map>
#include <stdexcept>
#include <string>
#include <vector>

namespace nlohmann {

class json {
public:
    enum class kind { null, number, string, array, object, boolean };

    json() : k_(kind::null), num_(0), is_int_(false), b_(false) {}
    json(double v) : k_(kind::number), num_(v), is_int_(false), b_(false) {}
    json(float v) : json(static_cast<double>(v)) {}
    json(int v) : k_(kind::number), num_(v), is_int_(true), b_(false) {}
    json(long v) : k_(kind::number), num_(static_cast<double>(v)), is_int_(true), b_(false) {}
    json(bool v) : k_(kind::boolean), num_(0), is_int_(false), b_(v) {}
    json(const char* s) : k_(kind::string), num_(0), is_int_(false), b_(false), str_(s) {}
    json(const std::string& s) : k_(kind::string), num_(0), is_int_(false), b_(false), str_(s) {}
    template <class T>
    json(const std::vector<T>& v) : k_(kind::array), num_(0), is_int_(false), b_(false) {
        for (const auto& x : v) arr_.emplace_back(x);
    }

    json& operator[](const std::string& key) {
        if (k_ == kind::null) k_ = kind::object;
        if (k_ != kind::object) throw std::runtime_error("json: not an object");
        return obj_[key];
    }
    const json& at(const std::string& key) const {
        auto it = obj_.find(key);
        if (k_ != kind::object || it == obj_.end()) throw std::out_of_range("json: key '" + key + "' not found");
        return it->second;
    }
    bool contains(const std::string& key) const { return k_ == kind::object && obj_.count(key) > 0; }
    size_t size() const { return k_ == kind::array ? arr_.size() : (k_ == kind::object ? obj_.size() : 0); }

    template <class T>
    T get() const {
        return get_impl(static_cast<T*>(nullptr));
    }

    std::string dump() const {
        std::string out;
        dump_to(out);
        return out;
    }

    static json parse(const std::string& text) {
        size_t pos = 0;
        json v = parse_value(text, pos);
        skip_ws(text, pos);
        if (pos != text.size()) throw std::runtime_error("json: trailing characters");
        return v;
    }

private:
    kind k_;
    double num_;
    bool is_int_, b_;
    std::string str_;
    std::vector<json> arr_;
    std::map<std::string, json> obj_;

    double number() const {
        if (k_ != kind::number) throw std::runtime_error("json: not a number");
        return num_;
    }
    float get_impl(float*) const { return static_cast<float>(number()); }
    double get_impl(double*) const { return number(); }
    int get_impl(int*) const { return static_cast<int>(number()); }
    long get_impl(long*) const { return static_cast<long>(number()); }
    bool get_impl(bool*) const { return k_ == kind::boolean ? b_ : number() != 0; }
    std::string get_impl(std::string*) const {
        if (k_ != kind::string) throw std::runtime_error("json: not a string");
        return str_;
    }
    template <class T>
    std::vector<T> get_impl(std::vector<T>*) const {
        if (k_ != kind::array) throw std::runtime_error("json: not an array");
        std::vector<T> v;
        for (const auto& x : arr_) v.push_back(x.get<T>());
        return v;
    }

    static void dump_string(const std::string& s, std::string& out) {
        out.push_back('"');
        for (char c : s) {
            if (c == '"' || c == '\\') { out.push_back('\\'); out.push_back(c); }
            else if (c == '\n') out += "\\n";
            else if (c == '\t') out += "\\t";
            else out.push_back(c);
        }
        out.push_back('"');
    }
    void dump_to(std::string& out) const {
        char buf[64];
        switch (k_) {
            case kind::null: out += "null"; break;
            case kind::boolean: out += b_ ? "true" : "false"; break;
            case kind::number:
                if (is_int_) snprintf(buf, sizeof(buf), "%ld", static_cast<long>(num_));
                else if (std::isfinite(num_)) snprintf(buf, sizeof(buf), "%.17g", num_);
                else snprintf(buf, sizeof(buf), "null");
                out += buf;
                break;
            case kind::string: dump_string(str_, out); break;
            case kind::array: {
                out.push_back('[');
                for (size_t i = 0; i < arr_.size(); ++i) {
                    if (i) out.push_back(',');
                    arr_[i].dump_to(out);
                }
                out.push_back(']');
                break;
            }
            case kind::object: {
                out.push_back('{');
                bool first = true;
                for (const auto& kv : obj_) {
                    if (!first) out.push_back(',');
                    first = false;
                    dump_string(kv.first, out);
                    out.push_back(':');
                    kv.second.dump_to(out);
                }
                out.push_back('}');
                break;
            }
        }
    }

    static void skip_ws(const std::string& t, size_t& p) {
        while (p < t.size() && (t[p] == ' ' || t[p] == '\n' || t[p] == '\t' || t[p] == '\r')) ++p;
    }
    static std::string parse_string(const std::string& t, size_t& p) {
        std::string s;
        ++p;
        while (p < t.size() && t[p] != '"') {
            if (t[p] == '\\' && p + 1 < t.size()) {
                ++p;
                switch (t[p]) {
                    case 'n': s.push_back('\n'); break;
                    case 't': s.push_back('\t'); break;
                    case 'r': s.push_back('\r'); break;
                    default: s.push_back(t[p]); break;
                }
            } else {
                s.push_back(t[p]);
            }
            ++p;
        }
        if (p >= t.size()) throw std::runtime_error("json: unterminated string");
        ++p;
        return s;
    }
    static json parse_value(const std::string& t, size_t& p) {
        skip_ws(t, p);
        if (p >= t.size()) throw std::runtime_error("json: unexpected end");
        const char c = t[p];
        if (c == '{') {
            json o;
            o.k_ = kind::object;
            ++p;
            skip_ws(t, p);
            if (p < t.size() && t[p] == '}') { ++p; return o; }
            for (;;) {
                skip_ws(t, p);
                if (p >= t.size() || t[p] != '"') throw std::runtime_error("json: key expected");
                const std::string key = parse_string(t, p);
                skip_ws(t, p);
                if (p >= t.size() || t[p] != ':') throw std::runtime_error("json: ':' expected");
                ++p;
                o.obj_[key] = parse_value(t, p);
                skip_ws(t, p);
                if (p < t.size() && t[p] == ',') { ++p; continue; }
                if (p < t.size() && t[p] == '}') { ++p; return o; }
                throw std::runtime_error("json: ',' or '}' expected");
            }
        }
        if (c == '[') {
            json a;
            a.k_ = kind::array;
            ++p;
            skip_ws(t, p);
            if (p < t.size() && t[p] == ']') { ++p; return a; }
            for (;;) {
                a.arr_.push_back(parse_value(t, p));
                skip_ws(t, p);
                if (p < t.size() && t[p] == ',') { ++p; continue; }
                if (p < t.size() && t[p] == ']') { ++p; return a; }
                throw std::runtime_error("json: ',' or ']' expected");
            }
        }
        if (c == '"') return json(parse_string(t, p));
        if (t.compare(p, 4, "true") == 0) { p += 4; return json(true); }
        if (t.compare(p, 5, "false") == 0) { p += 5; return json(false); }
        if (t.compare(p, 4, "null") == 0) { p += 4; return json(); }
        char* end = nullptr;
        const double v = std::strtod(t.c_str() + p, &end);
        if (end == t.c_str() + p) throw std::runtime_error("json: value expected");
        bool is_int = true;
        for (const char* q = t.c_str() + p; q < end; ++q)
            if (*q == '.' || *q == 'e' || *q == 'E') is_int = false;
        p = static_cast<size_t>(end - t.c_str());
        json n(v);
        n.is_int_ = is_int;
        return n;
    }
};

}  // namespace nlohmann

#endif  // INCLUDE_NLOHMANN_JSON_HPP_
#endif
