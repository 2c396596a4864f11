// Minimal boost::property_tree stand-in (TEST INFRASTRUCTURE): only what
// KalmanFilter::loadConfigurationFiles uses (KF.cpp:757-883): read_xml, get_child,
// iteration over children, get<T>("<xmlattr>.name", default).
#pragma once
#include <cctype>
#include <cstdlib>
#include <istream>
#include <iterator>
#include <sstream>
#include <stdexcept>
#include <string>
#include <utility>
#include <vector>

namespace boost {
class exception {
public:
    virtual ~exception() {}
};
namespace property_tree {

class ptree_error : public boost::exception, public std::runtime_error {
public:
    explicit ptree_error(const std::string &w) : std::runtime_error(w) {}
};

class ptree {
public:
    typedef std::pair<std::string, ptree> value_type;
    typedef std::vector<value_type>::const_iterator const_iterator;
    std::string data;
    std::vector<value_type> children;

    const_iterator begin() const { return children.begin(); }
    const_iterator end() const { return children.end(); }

    const ptree *find_path(const std::string &path) const {
        const ptree *cur = this;
        std::size_t pos = 0;
        while (pos <= path.size()) {
            std::size_t dot = path.find('.', pos);
            std::string key = path.substr(pos, dot == std::string::npos ? std::string::npos : dot - pos);
            const ptree *next = 0;
            for (std::size_t i = 0; i < cur->children.size(); ++i)
                if (cur->children[i].first == key) { next = &cur->children[i].second; break; }
            if (!next) return 0;
            cur = next;
            if (dot == std::string::npos) break;
            pos = dot + 1;
        }
        return cur;
    }
    const ptree &get_child(const std::string &path) const {
        const ptree *p = find_path(path);
        if (!p) throw ptree_error("No such node (" + path + ")");
        return *p;
    }
    template <typename T>
    T get(const std::string &path, const T &def) const {
        const ptree *p = find_path(path);
        if (!p) return def;
        std::istringstream ss(p->data);
        T v;
        ss >> v;
        if (ss.fail()) return def;
        ss >> std::ws;
        if (!ss.eof()) return def; // lexical conversion must consume everything
        return v;
    }
};

namespace detail {
inline void skip_ws(const std::string &s, std::size_t &i) {
    while (i < s.size() && std::isspace((unsigned char)s[i])) ++i;
}
inline bool name_char(char c) { return std::isalnum((unsigned char)c) || c == '_' || c == '-' || c == ':' || c == '.'; }

inline void parse_nodes(const std::string &s, std::size_t &i, ptree &parent, const std::string &closing) {
    for (;;) {
        std::size_t lt = s.find('<', i);
        if (lt == std::string::npos) {
            if (!closing.empty()) throw ptree_error("unexpected end of XML");
            i = s.size();
            return;
        }
        i = lt;
        if (s.compare(i, 4, "<!--") == 0) {
            std::size_t e = s.find("-->", i + 4);
            if (e == std::string::npos) throw ptree_error("unterminated comment");
            i = e + 3;
            continue;
        }
        if (s.compare(i, 2, "<?") == 0) {
            std::size_t e = s.find("?>", i + 2);
            if (e == std::string::npos) throw ptree_error("unterminated declaration");
            i = e + 2;
            continue;
        }
        if (s.compare(i, 2, "</") == 0) {
            std::size_t e = s.find('>', i);
            if (e == std::string::npos) throw ptree_error("unterminated closing tag");
            std::string nm = s.substr(i + 2, e - i - 2);
            while (!nm.empty() && std::isspace((unsigned char)nm[nm.size() - 1])) nm.erase(nm.size() - 1);
            if (nm != closing) throw ptree_error("mismatched closing tag");
            i = e + 1;
            return;
        }
        ++i;
        std::size_t n0 = i;
        while (i < s.size() && name_char(s[i])) ++i;
        if (i == n0) throw ptree_error("bad tag name");
        std::string name = s.substr(n0, i - n0);
        ptree node, attrs;
        bool self_closed = false;
        for (;;) {
            skip_ws(s, i);
            if (i >= s.size()) throw ptree_error("unterminated tag");
            if (s[i] == '/') {
                if (i + 1 >= s.size() || s[i + 1] != '>') throw ptree_error("bad tag end");
                i += 2;
                self_closed = true;
                break;
            }
            if (s[i] == '>') { ++i; break; }
            std::size_t a0 = i;
            while (i < s.size() && name_char(s[i])) ++i;
            if (i == a0) throw ptree_error("bad attribute");
            std::string an = s.substr(a0, i - a0);
            skip_ws(s, i);
            if (i >= s.size() || s[i] != '=') throw ptree_error("attribute without value");
            ++i;
            skip_ws(s, i);
            if (i >= s.size() || (s[i] != '"' && s[i] != '\'')) throw ptree_error("unquoted attribute");
            char q = s[i++];
            std::size_t v0 = i;
            while (i < s.size() && s[i] != q) ++i;
            if (i >= s.size()) throw ptree_error("unterminated attribute");
            ptree val;
            val.data = s.substr(v0, i - v0);
            attrs.children.push_back(std::make_pair(an, val));
            ++i;
        }
        if (!attrs.children.empty()) node.children.push_back(std::make_pair(std::string("<xmlattr>"), attrs));
        if (!self_closed) parse_nodes(s, i, node, name);
        parent.children.push_back(std::make_pair(name, node));
    }
}
} // namespace detail

template <typename Stream>
inline void read_xml(Stream &in, ptree &tree) {
    std::string s((std::istreambuf_iterator<char>(in)), std::istreambuf_iterator<char>());
    std::size_t i = 0;
    tree = ptree();
    detail::parse_nodes(s, i, tree, "");
    if (tree.children.empty()) throw ptree_error("no element found");
}

} // namespace property_tree
} // namespace boost
