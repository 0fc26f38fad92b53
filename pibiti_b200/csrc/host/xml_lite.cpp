#include "xml_lite.h"
#include <cstdio>
#include <cstring>
#include <cstdlib>

namespace sphxml {

const char* Element::Attribute(const char* key) const
{
    for (const auto& kv : attrs)
        if (kv.first == key) return kv.second.c_str();
    return nullptr;
}

const Element* Element::FirstChildElement(const char* childName) const
{
    for (const auto& c : children)
        if (c->name == childName) return c.get();
    return nullptr;
}

const Element* Element::NextSiblingElement(const char* siblingName) const
{
    if (!parent) return nullptr;
    bool seen = false;
    for (const auto& c : parent->children) {
        if (seen && c->name == siblingName) return c.get();
        if (c.get() == this) seen = true;
    }
    return nullptr;
}

namespace {

struct Cursor {
    const std::string& s;
    size_t i = 0;
    explicit Cursor(const std::string& t) : s(t) {}
    bool eof() const { return i >= s.size(); }
    bool starts(const char* lit) const { return s.compare(i, strlen(lit), lit) == 0; }
    void skip_ws() { while (!eof() && (unsigned char)s[i] <= ' ') i++; }
};

bool is_name_char(char c)
{
    return (c >= 'a' && c <= 'z') || (c >= 'A' && c <= 'Z') || (c >= '0' && c <= '9') || c == '_' || c == '-' ||
           c == ':' || c == '.' || (unsigned char)c >= 0x80;
}

std::string decode_entities(const std::string& raw)
{
    std::string out;
    out.reserve(raw.size());
    for (size_t i = 0; i < raw.size(); i++) {
        if (raw[i] != '&') { out += raw[i]; continue; }
        size_t semi = raw.find(';', i);
        if (semi == std::string::npos || semi - i > 10) { out += raw[i]; continue; }
        std::string ent = raw.substr(i + 1, semi - i - 1);
        if (ent == "amp") out += '&';
        else if (ent == "lt") out += '<';
        else if (ent == "gt") out += '>';
        else if (ent == "quot") out += '"';
        else if (ent == "apos") out += '\'';
        else if (!ent.empty() && ent[0] == '#') {
            long v = ent.size() > 1 && (ent[1] == 'x' || ent[1] == 'X') ? strtol(ent.c_str() + 2, nullptr, 16)
                                                                          : strtol(ent.c_str() + 1, nullptr, 10);
            if (v > 0 && v < 128) out += (char)v;
        } else { out += raw[i]; continue; }
        i = semi;
    }
    return out;
}

// skips comments, processing instructions, DOCTYPE and text up to the next '<' that opens or
// closes an element; returns false at end of input
bool seek_tag(Cursor& c)
{
    for (;;) {
        size_t lt = c.s.find('<', c.i);
        if (lt == std::string::npos) { c.i = c.s.size(); return false; }
        c.i = lt;
        if (c.starts("<!--")) {
            size_t end = c.s.find("-->", c.i + 4);
            c.i = end == std::string::npos ? c.s.size() : end + 3;
        } else if (c.starts("<?")) {
            size_t end = c.s.find("?>", c.i + 2);
            c.i = end == std::string::npos ? c.s.size() : end + 2;
        } else if (c.starts("<![CDATA[")) {
            size_t end = c.s.find("]]>", c.i + 9);
            c.i = end == std::string::npos ? c.s.size() : end + 3;
        } else if (c.starts("<!")) {
            size_t end = c.s.find('>', c.i + 2);
            c.i = end == std::string::npos ? c.s.size() : end + 1;
        } else {
            return true;
        }
    }
}

bool parse_element(Cursor& c, Element* parent, std::unique_ptr<Element>& out, std::string& err)
{
    // c.i at '<' of an opening tag
    c.i++;
    size_t n0 = c.i;
    while (!c.eof() && is_name_char(c.s[c.i])) c.i++;
    if (c.i == n0) { err = "element name expected at offset " + std::to_string(n0); return false; }
    out.reset(new Element());
    out->name = c.s.substr(n0, c.i - n0);
    out->parent = parent;

    for (;;) {                                            // attributes
        c.skip_ws();
        if (c.eof()) { err = "unterminated tag <" + out->name; return false; }
        if (c.starts("/>")) { c.i += 2; return true; }
        if (c.s[c.i] == '>') { c.i++; break; }
        size_t a0 = c.i;
        while (!c.eof() && is_name_char(c.s[c.i])) c.i++;
        if (c.i == a0) { err = "attribute name expected in <" + out->name + " at offset " + std::to_string(a0); return false; }
        std::string key = c.s.substr(a0, c.i - a0);
        c.skip_ws();
        if (c.eof() || c.s[c.i] != '=') { err = "'=' expected after attribute " + key; return false; }
        c.i++;
        c.skip_ws();
        if (c.eof() || (c.s[c.i] != '"' && c.s[c.i] != '\'')) { err = "quoted value expected for attribute " + key; return false; }
        char q = c.s[c.i++];
        size_t v0 = c.i;
        size_t v1 = c.s.find(q, v0);
        if (v1 == std::string::npos) { err = "unterminated value of attribute " + key; return false; }
        out->attrs.emplace_back(key, decode_entities(c.s.substr(v0, v1 - v0)));
        c.i = v1 + 1;
    }

    for (;;) {                                            // content
        if (!seek_tag(c)) { err = "missing </" + out->name + ">"; return false; }
        if (c.starts("</")) {
            size_t end = c.s.find('>', c.i);
            if (end == std::string::npos) { err = "unterminated closing tag"; return false; }
            c.i = end + 1;
            return true;
        }
        std::unique_ptr<Element> child;
        if (!parse_element(c, out.get(), child, err)) return false;
        out->children.push_back(std::move(child));
    }
}

}  // namespace

bool Document::Parse(const std::string& text)
{
    root.reset();
    error.clear();
    Cursor c(text);
    if (!seek_tag(c) || c.starts("</")) { error = "no root element"; return false; }
    return parse_element(c, nullptr, root, error);
}

bool Document::LoadFile(const char* path)
{
    root.reset();
    FILE* f = fopen(path, "rb");
    if (!f) { error = std::string("cannot open ") + path; return false; }
    std::string text;
    char buf[65536];
    size_t got;
    while ((got = fread(buf, 1, sizeof buf, f)) > 0) text.append(buf, got);
    fclose(f);
    return Parse(text);
}

}  // namespace sphxml
