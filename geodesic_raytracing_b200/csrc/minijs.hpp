// A small JavaScript interpreter - just enough of ES5/ES6 to run the reference's metric plugin
// scripts unchanged (scripts/*.js, scripts/coordinates/*.js, scripts/origins/*.js).
//
// The reference embeds QuickJS (2020-04-12, CONFIG_BIGNUM) and gets `a * b` on `number` objects
// through `Operators.create` in /root/reference/number.js:4-71.  This image has no JavaScript
// engine and the reference's vendored one must not be copied, so the plugin boundary
//  is served by this interpreter: operator overloading on host objects is a
// first-class hook instead of the bignum extension, everything else is plain JavaScript:
// var/let/const with hoisting, functions/closures/arrows, default and rest parameters, destructuring declarations and
// parameters, spread in calls / array / object literals, template literals, if/for/for-in/for-of/while/do/switch, try/catch,
// classes (constructor, methods, static methods, getters, extends + super), arrays with the usual methods, objects, Math.*,
// the usual operators, automatic semicolon insertion.  What is not implemented (generators, async, regular expressions,
// labels, destructuring assignment without a declaration, class fields, tagged templates, optional chaining) is refused with
// a SyntaxError that names the script and the line - never skipped silently.  The completion value of a
// script (the value of its last expression statement) is what `eval` returns, which is how the
// reference obtains the metric function (/root/reference/js_interop.cpp:738-742).
#pragma once
#include <functional>
#include <algorithm>
#include <map>
#include <memory>
#include <stdexcept>
#include <set>
#include <string>
#include <unordered_map>
#include <vector>

namespace minijs {

struct Object;
struct Interp;
using ObjP = std::shared_ptr<Object>;

struct Value {
    enum Type { Undefined, Null, Bool, Number, String, Obj } type = Undefined;
    double num = 0;
    bool b = false;
    std::string str;
    ObjP obj;
    Value() {}
    Value(double d) : type(Number), num(d) {}
    Value(int d) : type(Number), num(d) {}
    Value(bool v) : type(Bool), b(v) {}
    Value(const std::string& s) : type(String), str(s) {}
    Value(const char* s) : type(String), str(s) {}
    Value(ObjP o) : type(Obj), obj(std::move(o)) {}
    static Value null() { Value v; v.type = Null; return v; }
    bool is_undefined() const { return type == Undefined; }
    bool is_number() const { return type == Number; }
    bool is_bool() const { return type == Bool; }
    bool is_object() const { return type == Obj; }
};

struct JsError : std::runtime_error {
    using std::runtime_error::runtime_error;
};

using NativeFn = std::function<Value(Interp&, const Value& self, std::vector<Value>& args)>;

struct FuncDef;  // AST
struct Env;
using EnvP = std::shared_ptr<Env>;

// The own properties of an object in JavaScript's enumeration order: integer-like keys ascending, then the other keys in the order they
// were created (for-in, Object.keys, spread and object rest see this order; the reference's QuickJS does the same).  The interface is
// the part of std::map the interpreter uses.  Objects are small: lookups go through a hash index, the order is rebuilt lazily.
class PropMap {
public:
    using value_type = std::pair<std::string, Value>;
    template <class Map, class Ref>
    class Iter {
    public:
        Iter(Map* m, size_t pos) : m_(m), pos_(pos) {}
        Ref& operator*() const { return m_->items_[m_->order_[pos_]]; }
        Ref* operator->() const { return &m_->items_[m_->order_[pos_]]; }
        Iter& operator++() { ++pos_; return *this; }
        bool operator==(const Iter& o) const { return pos_ == o.pos_; }
        bool operator!=(const Iter& o) const { return pos_ != o.pos_; }
    private:
        Map* m_;
        size_t pos_;
    };
    using iterator = Iter<PropMap, value_type>;
    using const_iterator = Iter<const PropMap, const value_type>;

    Value& operator[](const std::string& key) {
        auto it = index_.find(key);
        if (it != index_.end()) return items_[it->second].second;
        index_.emplace(key, items_.size());
        items_.emplace_back(key, Value());
        dirty_ = true;
        return items_.back().second;
    }
    iterator find(const std::string& key) { return iterator(this, position(key)); }
    const_iterator find(const std::string& key) const { return const_iterator(this, position(key)); }
    size_t count(const std::string& key) const { return index_.count(key); }
    size_t size() const { return items_.size(); }
    bool empty() const { return items_.empty(); }
    void clear() { items_.clear(); index_.clear(); order_.clear(); dirty_ = false; }
    bool erase(const std::string& key) {
        auto it = index_.find(key);
        if (it == index_.end()) return false;
        const size_t at = it->second;
        items_.erase(items_.begin() + (long)at);
        index_.erase(it);
        for (auto& kv : index_) if (kv.second > at) kv.second--;
        dirty_ = true;
        return true;
    }
    iterator begin() { sort(); return iterator(this, 0); }
    iterator end() { return iterator(this, items_.size()); }
    const_iterator begin() const { sort(); return const_iterator(this, 0); }
    const_iterator end() const { return const_iterator(this, items_.size()); }

private:
    static bool array_index(const std::string& k, unsigned long long& out) {
        if (k.empty() || k.size() > 10 || (k.size() > 1 && k[0] == '0')) return false;
        out = 0;
        for (char c : k) { if (c < '0' || c > '9') return false; out = out * 10 + (unsigned long long)(c - '0'); }
        return out < 4294967295ull;
    }
    void sort() const {
        if (!dirty_ && order_.size() == items_.size()) return;
        std::vector<std::pair<unsigned long long, size_t>> ints;
        std::vector<size_t> rest;
        for (size_t i = 0; i < items_.size(); i++) {
            unsigned long long v;
            if (array_index(items_[i].first, v)) ints.emplace_back(v, i); else rest.push_back(i);
        }
        std::sort(ints.begin(), ints.end());
        order_.clear();
        for (auto& p : ints) order_.push_back(p.second);
        order_.insert(order_.end(), rest.begin(), rest.end());
        dirty_ = false;
    }
    size_t position(const std::string& key) const {
        auto it = index_.find(key);
        if (it == index_.end()) return items_.size();
        sort();
        for (size_t i = 0; i < order_.size(); i++) if (order_[i] == it->second) return i;
        return items_.size();
    }
    std::vector<value_type> items_;                       // creation order
    std::unordered_map<std::string, size_t> index_;
    mutable std::vector<size_t> order_;                   // enumeration order: indices into items_
    mutable bool dirty_ = false;
    template <class, class> friend class Iter;
};

struct Object {
    enum Kind { Plain, Array, Function, Native, Host } kind = Plain;
    PropMap props;
    std::vector<Value> arr;                 // Array
    const FuncDef* def = nullptr;           // Function
    EnvP env;                               // Function closure
    Value bound_this;                       // arrow functions capture `this`
    bool is_arrow = false;
    NativeFn native;                        // Native
    int host_tag = 0;                       // Host: type tag chosen by the embedder
    std::shared_ptr<void> host;             // Host payload
    ObjP klass;                             // an error object: its kind (a native constructor); a class / error kind: its parent
    ObjP proto;                             // the prototype: where a property that is not the object's own is looked up (instances -> C.prototype -> B.prototype ...; a class -> its parent class, for statics)
    bool class_proto = false;               // a prototype object made by `class`: its members do not show in for-in
    std::set<std::string> hidden;           // own properties that do not enumerate (Object.defineProperty without enumerable: true)
    // Map / Set: entries in insertion order (a Set keeps its values as keys); `is_collection` 1 = Map, 2 = Set
    int is_collection = 0;
    std::vector<std::pair<Value, Value>> entries;
    bool is_class = false;                  // a constructor made by `class`: props["prototype"] holds its methods
    // property hooks for Host objects (return true when handled)
    std::function<bool(Interp&, const std::string&, Value&)> get_hook;
    std::function<bool(Interp&, const std::string&, const Value&)> set_hook;
};

struct Env {
    std::unordered_map<std::string, Value> vars;
    std::unordered_map<std::string, bool> is_const;
    EnvP parent;
    bool function_scope = false;
    Value this_val;
};

struct Program;  // parsed script (owns its AST)

struct Interp {
    Interp();
    ~Interp();
    // Parse + run; returns the completion value. Throws JsError (message carries the line).
    Value eval(const std::string& source, const std::string& name = "<script>");
    Value call(const Value& fn, const Value& self, std::vector<Value> args);
    void set_global(const std::string& name, const Value& v);
    Value get_global(const std::string& name);
    // global with a getter (the reference installs `$cfg` this way, js_interop.cpp:841-846)
    void set_global_getter(const std::string& name, std::function<Value(Interp&)> getter);

    static ObjP make_object();
    static ObjP make_array(std::vector<Value> items = {});
    static ObjP make_native(NativeFn fn);
    static ObjP make_host(int tag, std::shared_ptr<void> payload);

    // operator overloading: called when either operand is a Host object.
    // op is one of "+", "-", "*", "/", "%", "**", "neg", "pos", "<", "<=", ">", ">=", "==", "!=".
    std::function<Value(Interp&, const std::string& op, const Value& a, const Value& b)> host_operator;

    // conversions (ECMAScript semantics for the primitive cases)
    static double to_number(const Value& v);
    static bool to_boolean(const Value& v);
    std::string to_string(const Value& v);

    EnvP globals;
    ObjP builtin_array, builtin_object, builtin_function;      // the constructors `instanceof` knows without a class chain
    std::map<std::string, ObjP> error_types;                  // Error, TypeError, RangeError, ReferenceError, SyntaxError
    // the object a `catch` binds for an error raised by the interpreter itself ("... TypeError: cannot read ..."): message, name, class
    Value make_error_object(const std::string& what);
    static std::vector<std::string> own_keys(const Value& v);   // enumerable own property names, JavaScript's order
    // what for-of, spread and destructuring walk over: the elements of an array, the characters of a string, the [key, value] pairs of
    // a Map, the values of a Set; false if the value is not iterable
    static bool iterate(const Value& v, std::vector<Value>& out);
    std::vector<std::unique_ptr<Program>> programs;  // keep ASTs alive for closures
    std::unordered_map<std::string, std::function<Value(Interp&)>> global_getters;
    int call_depth = 0;
    long steps = 0;  // runaway-script guard
};

}  // namespace minijs
