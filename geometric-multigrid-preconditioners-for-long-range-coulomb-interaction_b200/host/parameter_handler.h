// A small stand-in for dealii::ParameterHandler / Patterns with the subset of the interface the
// reference uses (src/main.cc:20-59, src/step-50.cc:13-101): declare_entry, enter/leave_subsection,
// parse_input (file), parse_input_from_string, get / get_integer / get_double / get_bool.
// Same `.prm` grammar: `subsection NAME` ... `end`, `set KEY = VALUE`, `#` comments.
#pragma once
#include <fstream>
#include <map>
#include <memory>
#include <sstream>
#include <stdexcept>
#include <string>
#include <vector>

namespace Patterns {
struct PatternBase {
  virtual ~PatternBase() = default;
  virtual bool match(const std::string &v) const = 0;
  virtual std::string description() const = 0;
};
struct Integer : PatternBase {
  bool match(const std::string &v) const override {
    if (v.empty()) return false;
    size_t pos = 0;
    try { (void)std::stol(v, &pos); } catch (...) { return false; }
    return pos == v.size();
  }
  std::string description() const override { return "[Integer]"; }
};
struct Double : PatternBase {
  bool match(const std::string &v) const override {
    if (v.empty()) return false;
    size_t pos = 0;
    try { (void)std::stod(v, &pos); } catch (...) { return false; }
    return pos == v.size();
  }
  std::string description() const override { return "[Double]"; }
};
struct Bool : PatternBase {
  bool match(const std::string &v) const override {
    return v == "true" || v == "false" || v == "yes" || v == "no" || v == "on" || v == "off";
  }
  std::string description() const override { return "[Bool]"; }
};
struct Anything : PatternBase {
  bool match(const std::string &) const override { return true; }
  std::string description() const override { return "[Anything]"; }
};
struct Selection : PatternBase {
  std::vector<std::string> options;
  explicit Selection(const std::string &seq) {
    std::string cur;
    auto flush = [&]() {
      size_t a = cur.find_first_not_of(" \t"), b = cur.find_last_not_of(" \t");
      if (a != std::string::npos) options.push_back(cur.substr(a, b - a + 1));
      cur.clear();
    };
    for (char c : seq) {
      if (c == '|') flush();
      else cur += c;
    }
    flush();
  }
  bool match(const std::string &v) const override {
    for (auto &o : options)
      if (o == v) return true;
    return false;
  }
  std::string description() const override {
    std::string s = "[Selection ";
    for (size_t i = 0; i < options.size(); ++i) s += (i ? "|" : "") + options[i];
    return s + " ]";
  }
};
}  // namespace Patterns

class ExcParameter : public std::runtime_error {
 public:
  explicit ExcParameter(const std::string &m) : std::runtime_error(m) {}
};

class ParameterHandler {
 public:
  template <class P>
  void declare_entry(const std::string &entry, const std::string &default_value, const P &pattern,
                     const std::string &documentation = "") {
    Entry e;
    e.value = e.default_value = default_value;
    e.pattern = std::make_shared<P>(pattern);
    e.doc = documentation;
    if (!e.pattern->match(default_value))
      throw ExcParameter("default value <" + default_value + "> of entry <" + entry + "> does not match its pattern");
    entries[path_key(entry)] = e;
  }
  void enter_subsection(const std::string &s) { path.push_back(s); }
  void leave_subsection() {
    if (path.empty()) throw ExcParameter("leave_subsection without enter_subsection");
    path.pop_back();
  }
  void parse_input(const std::string &filename) {
    std::ifstream f(filename);
    if (!f.is_open()) throw ExcParameter("could not open parameter file <" + filename + ">");
    std::stringstream ss;
    ss << f.rdbuf();
    parse_input_from_string(ss.str().c_str());
  }
  void parse_input_from_string(const char *s) {
    std::istringstream in(s);
    std::string raw;
    const std::vector<std::string> saved = path;
    int lineno = 0;
    while (std::getline(in, raw)) {
      ++lineno;
      std::string line = raw.substr(0, raw.find('#'));
      line = trim(line);
      if (line.empty()) continue;
      std::string lower = line;
      for (auto &c : lower) c = (char)std::tolower(c);
      if (lower.rfind("subsection", 0) == 0 && (line.size() == 10 || std::isspace((unsigned char)line[10]))) {
        path.push_back(squeeze(line.substr(10)));
      } else if (lower == "end") {
        if (path.size() <= saved.size()) throw ExcParameter("line " + std::to_string(lineno) + ": unbalanced 'end'");
        path.pop_back();
      } else if (lower.rfind("set", 0) == 0 && line.size() > 3 && std::isspace((unsigned char)line[3])) {
        const size_t eq = line.find('=');
        if (eq == std::string::npos) throw ExcParameter("line " + std::to_string(lineno) + ": missing '='");
        const std::string name = squeeze(line.substr(3, eq - 3));
        const std::string value = trim(line.substr(eq + 1));
        auto it = entries.find(path_key(name));
        if (it == entries.end())
          throw ExcParameter("line " + std::to_string(lineno) + ": No entry with name <" + name +
                             "> was declared in the current subsection.");
        if (!it->second.pattern->match(value))
          throw ExcParameter("line " + std::to_string(lineno) + ": The entry value <" + value + "> for the entry named <" +
                             name + "> does not match the given pattern " + it->second.pattern->description());
        it->second.value = value;
      } else {
        throw ExcParameter("line " + std::to_string(lineno) + ": could not parse <" + raw + ">");
      }
    }
    if (path.size() != saved.size()) {
      path = saved;
      throw ExcParameter("unbalanced 'subsection'/'end' in input");
    }
  }
  std::string get(const std::string &entry) const {
    auto it = entries.find(path_key(entry));
    if (it == entries.end()) throw ExcParameter("You can't ask for entry <" + entry + "> you have not yet declared.");
    return it->second.value;
  }
  long get_integer(const std::string &entry) const { return std::stol(get(entry)); }
  double get_double(const std::string &entry) const { return std::stod(get(entry)); }
  bool get_bool(const std::string &entry) const {
    const std::string v = get(entry);
    return v == "true" || v == "yes" || v == "on";
  }

 private:
  struct Entry {
    std::string value, default_value, doc;
    std::shared_ptr<Patterns::PatternBase> pattern;
  };
  std::map<std::string, Entry> entries;
  std::vector<std::string> path;
  static std::string trim(const std::string &s) {
    const size_t a = s.find_first_not_of(" \t\r\n"), b = s.find_last_not_of(" \t\r\n");
    return a == std::string::npos ? "" : s.substr(a, b - a + 1);
  }
  static std::string squeeze(const std::string &s) {  // trim + collapse inner whitespace
    std::istringstream in(s);
    std::string w, out;
    while (in >> w) out += (out.empty() ? "" : " ") + w;
    return out;
  }
  std::string path_key(const std::string &entry) const {
    std::string k;
    for (auto &p : path) k += p + "/";
    return k + entry;
  }
};
