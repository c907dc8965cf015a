// TEST INFRASTRUCTURE — not part of the product path.
//
// Header-only stand-in for <boost/program_options.hpp> with the surface the
// reference CLI uses (ribbit.cpp:82-112): options_description,
// add_options()(name,desc)(name,value<T>(),desc), variables_map (count,
// operator[], as<T>), parse_command_line, store, notify, operator<<.
// Accepts "--long value", "--long=value", "-s value" and "-svalue".
#ifndef RB_SHIM_PROGRAM_OPTIONS_HPP
#define RB_SHIM_PROGRAM_OPTIONS_HPP

#include <algorithm>
#include <cctype>
#include <cmath>
#include <cstdlib>
#include <map>
#include <memory>
#include <ostream>
#include <sstream>
#include <stdexcept>
#include <string>
#include <vector>

namespace boost {
namespace program_options {

struct error : std::runtime_error {
    explicit error(const std::string &m) : std::runtime_error(m) {}
};

struct value_semantic {
    virtual ~value_semantic() {}
    virtual std::shared_ptr<void> parse(const std::string &s) const = 0;
};

template <typename T>
struct typed_value : value_semantic {
    std::shared_ptr<void> parse(const std::string &s) const override {
        std::istringstream is(s);
        auto p = std::make_shared<T>();
        is >> *p;
        if (is.fail()) throw error("the argument ('" + s + "') is invalid");
        return p;
    }
};
template <>
struct typed_value<std::string> : value_semantic {
    std::shared_ptr<void> parse(const std::string &s) const override { return std::make_shared<std::string>(s); }
};

template <typename T>
inline typed_value<T> *value() { return new typed_value<T>(); }

struct option_description {
    std::string long_name, short_name, description;
    std::shared_ptr<const value_semantic> semantic;  // null => flag
};

class options_description;

class options_description_easy_init {
    options_description *owner_;

public:
    explicit options_description_easy_init(options_description *o) : owner_(o) {}
    options_description_easy_init &operator()(const char *name, const char *description);
    options_description_easy_init &operator()(const char *name, const value_semantic *s, const char *description);
};

class options_description {
public:
    explicit options_description(const std::string &caption = "") : caption_(caption) {}
    options_description_easy_init add_options() { return options_description_easy_init(this); }
    void add(const char *name, const value_semantic *s, const char *description) {
        option_description d;
        std::string n(name);
        std::size_t c = n.find(',');
        d.long_name = n.substr(0, c);
        if (c != std::string::npos) d.short_name = n.substr(c + 1);
        d.description = description ? description : "";
        d.semantic.reset(s);
        options_.push_back(d);
    }
    const option_description *find_long(const std::string &n) const {
        for (auto &o : options_) if (o.long_name == n) return &o;
        return nullptr;
    }
    const option_description *find_short(const std::string &n) const {
        for (auto &o : options_) if (!o.short_name.empty() && o.short_name == n) return &o;
        return nullptr;
    }
    friend std::ostream &operator<<(std::ostream &os, const options_description &d) {
        os << d.caption_ << ":\n";
        for (auto &o : d.options_) {
            os << "  ";
            if (!o.short_name.empty()) os << "-" << o.short_name << " [ --" << o.long_name << " ]";
            else os << "--" << o.long_name;
            if (o.semantic) os << " arg";
            os << "\t" << o.description << "\n";
        }
        return os;
    }

private:
    std::string caption_;
    std::vector<option_description> options_;
};

inline options_description_easy_init &options_description_easy_init::operator()(const char *name, const char *description) {
    owner_->add(name, nullptr, description);
    return *this;
}
inline options_description_easy_init &options_description_easy_init::operator()(const char *name, const value_semantic *s, const char *description) {
    owner_->add(name, s, description);
    return *this;
}

struct parsed_option {
    std::string key;
    std::shared_ptr<void> value;
};
struct parsed_options {
    std::vector<parsed_option> options;
};

class variable_value {
    std::shared_ptr<void> v_;

public:
    variable_value() {}
    explicit variable_value(std::shared_ptr<void> v) : v_(v) {}
    template <typename T>
    const T &as() const {
        if (!v_) throw error("bad any_cast: empty option value");
        return *static_cast<const T *>(v_.get());
    }
    bool empty() const { return !v_; }
};

class variables_map : public std::map<std::string, variable_value> {
public:
    std::size_t count(const std::string &k) const { return std::map<std::string, variable_value>::count(k); }
    const variable_value &operator[](const std::string &k) const {
        static const variable_value empty_value;
        auto it = find(k);
        return it == end() ? empty_value : it->second;
    }
};

inline parsed_options parse_command_line(int argc, const char *const argv[], const options_description &desc) {
    parsed_options out;
    for (int i = 1; i < argc; ++i) {
        std::string tok(argv[i]);
        const option_description *od = nullptr;
        std::string attached;
        bool has_attached = false;
        if (tok.size() > 2 && tok[0] == '-' && tok[1] == '-') {
            std::string name = tok.substr(2);
            std::size_t eq = name.find('=');
            if (eq != std::string::npos) { attached = name.substr(eq + 1); name = name.substr(0, eq); has_attached = true; }
            od = desc.find_long(name);
            if (!od) throw error("unrecognised option '" + tok + "'");
        } else if (tok.size() >= 2 && tok[0] == '-') {
            od = desc.find_short(tok.substr(1, 1));
            if (!od) throw error("unrecognised option '" + tok + "'");
            if (tok.size() > 2) { attached = tok.substr(2); has_attached = true; }
        } else {
            throw error("too many positional options have been specified on the command line");
        }
        parsed_option po;
        po.key = od->long_name;
        if (od->semantic) {
            std::string val;
            if (has_attached) val = attached;
            else {
                if (i + 1 >= argc) throw error("the required argument for option '--" + od->long_name + "' is missing");
                val = argv[++i];
            }
            po.value = od->semantic->parse(val);
        } else {
            po.value = std::make_shared<bool>(true);
        }
        out.options.push_back(po);
    }
    return out;
}

inline void store(const parsed_options &p, variables_map &vm) {
    for (auto &o : p.options)
        if (!vm.count(o.key)) vm.insert(std::make_pair(o.key, variable_value(o.value)));
}
inline void notify(variables_map &) {}

}  // namespace program_options
}  // namespace boost

#endif
