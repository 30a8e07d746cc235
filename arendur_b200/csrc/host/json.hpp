// Minimal JSON reader for arendur scene descriptions (examples/arencli.rs:206-215 uses
// serde_json).  Product host code; no external dependency.
#pragma once
#include <cstdlib>
#include <map>
#include <memory>
#include <string>
#include <vector>

namespace arnhost {

struct Json {
    enum Kind { Null, Bool, Num, Str, Arr, Obj } kind = Null;
    bool b = false; double num = 0.0; std::string str;
    std::vector<Json> arr;
    std::vector<std::pair<std::string, Json>> obj;   // keeps file order

    const Json* get(const std::string& key) const {
        if (kind != Obj) return nullptr;
        for (auto& kv : obj) if (kv.first == key) return &kv.second;
        return nullptr;
    }
    bool is_null() const { return kind == Null; }
};

class JsonParser {
public:
    explicit JsonParser(const std::string& text) : s(text), i(0) {}
    bool parse(Json* out, std::string* err) {
        if (!value(out)) { if (err) *err = msg + " at byte " + std::to_string(i); return false; }
        ws();
        if (i != s.size()) { if (err) *err = "trailing characters at byte " + std::to_string(i); return false; }
        return true;
    }
private:
    const std::string& s; size_t i; std::string msg;
    void ws() { while (i < s.size() && (s[i] == ' ' || s[i] == '\n' || s[i] == '\r' || s[i] == '\t')) i++; }
    bool fail(const char* m) { msg = m; return false; }
    bool value(Json* o) {
        ws();
        if (i >= s.size()) return fail("unexpected end");
        char c = s[i];
        if (c == '{') return object(o);
        if (c == '[') return array(o);
        if (c == '"') { o->kind = Json::Str; return string(&o->str); }
        if (s.compare(i, 4, "true") == 0) { o->kind = Json::Bool; o->b = true; i += 4; return true; }
        if (s.compare(i, 5, "false") == 0) { o->kind = Json::Bool; o->b = false; i += 5; return true; }
        if (s.compare(i, 4, "null") == 0) { o->kind = Json::Null; i += 4; return true; }
        return number(o);
    }
    bool number(Json* o) {
        const char* st = s.c_str() + i; char* en = nullptr;
        double v = std::strtod(st, &en);
        if (en == st) return fail("invalid value");
        o->kind = Json::Num; o->num = v; i += (size_t)(en - st); return true;
    }
    bool string(std::string* out) {
        i++; out->clear();
        while (i < s.size() && s[i] != '"') {
            if (s[i] == '\\' && i + 1 < s.size()) {
                char e = s[i + 1];
                switch (e) { case 'n': out->push_back('\n'); break; case 't': out->push_back('\t'); break; case 'r': out->push_back('\r'); break;
                             case 'b': out->push_back('\b'); break; case 'f': out->push_back('\f'); break;
                             case 'u': { if (i + 5 >= s.size()) return fail("bad \\u escape");
                                         unsigned cp = (unsigned)std::strtoul(s.substr(i + 2, 4).c_str(), nullptr, 16);
                                         if (cp < 0x80) out->push_back((char)cp);
                                         else if (cp < 0x800) { out->push_back((char)(0xC0 | (cp >> 6))); out->push_back((char)(0x80 | (cp & 0x3F))); }
                                         else { out->push_back((char)(0xE0 | (cp >> 12))); out->push_back((char)(0x80 | ((cp >> 6) & 0x3F))); out->push_back((char)(0x80 | (cp & 0x3F))); }
                                         i += 4; break; }
                             default: out->push_back(e); }
                i += 2;
            } else out->push_back(s[i++]);
        }
        if (i >= s.size()) return fail("unterminated string");
        i++; return true;
    }
    bool array(Json* o) {
        o->kind = Json::Arr; i++; ws();
        if (i < s.size() && s[i] == ']') { i++; return true; }
        for (;;) {
            Json v; if (!value(&v)) return false; o->arr.push_back(std::move(v)); ws();
            if (i < s.size() && s[i] == ',') { i++; continue; }
            if (i < s.size() && s[i] == ']') { i++; return true; }
            return fail("expected , or ]");
        }
    }
    bool object(Json* o) {
        o->kind = Json::Obj; i++; ws();
        if (i < s.size() && s[i] == '}') { i++; return true; }
        for (;;) {
            ws(); if (i >= s.size() || s[i] != '"') return fail("expected object key");
            std::string k; if (!string(&k)) return false; ws();
            if (i >= s.size() || s[i] != ':') return fail("expected :");
            i++;
            Json v; if (!value(&v)) return false; o->obj.emplace_back(std::move(k), std::move(v)); ws();
            if (i < s.size() && s[i] == ',') { i++; continue; }
            if (i < s.size() && s[i] == '}') { i++; return true; }
            return fail("expected , or }");
        }
    }
};

}  // namespace arnhost
