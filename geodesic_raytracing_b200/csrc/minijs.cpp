// Small JavaScript interpreter for the metric plugin scripts. See minijs.hpp.
#include "minijs.hpp"

#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <set>

namespace minijs {

// ------------------------------------------------------------------------------------------
// Lexer
// ------------------------------------------------------------------------------------------
enum class Tok { End, Num, Str, Ident, Keyword, Punct, Template };

struct Token {
    Tok type = Tok::End;
    std::string text;
    double num = 0;
    int line = 1;
    bool nl_before = false;
};

static const std::set<std::string> kKeywords = {
    "var", "let", "const", "function", "return", "if", "else", "for", "while", "do", "break",
    "continue", "new", "typeof", "this", "true", "false", "null", "undefined", "in", "of",
    "throw", "try", "catch", "finally", "void", "delete", "instanceof", "class", "switch",
    "case", "default"};

// one escape sequence after a backslash at src[pos] (pos is advanced): appends the character(s); code points above 0x7f as UTF-8
static void read_escape(const std::string& src, size_t& pos, std::string& out, int& line) {
    auto hex = [&](size_t count, unsigned& v) {
        v = 0;
        if (pos + count > src.size()) return false;
        for (size_t i = 0; i < count; i++) { if (!isxdigit((unsigned char)src[pos + i])) return false; v = v * 16 + (unsigned)(isdigit((unsigned char)src[pos + i]) ? src[pos + i] - '0' : tolower((unsigned char)src[pos + i]) - 'a' + 10); }
        pos += count;
        return true;
    };
    auto utf8 = [&](unsigned cp) {
        if (cp < 0x80) out += (char)cp;
        else if (cp < 0x800) { out += (char)(0xC0 | (cp >> 6)); out += (char)(0x80 | (cp & 0x3F)); }
        else if (cp < 0x10000) { out += (char)(0xE0 | (cp >> 12)); out += (char)(0x80 | ((cp >> 6) & 0x3F)); out += (char)(0x80 | (cp & 0x3F)); }
        else { out += (char)(0xF0 | (cp >> 18)); out += (char)(0x80 | ((cp >> 12) & 0x3F)); out += (char)(0x80 | ((cp >> 6) & 0x3F)); out += (char)(0x80 | (cp & 0x3F)); }
    };
    const char e = src[pos++];
    unsigned v = 0;
    switch (e) {
        case 'n': out += '\n'; break;
        case 't': out += '\t'; break;
        case 'r': out += '\r'; break;
        case 'b': out += '\b'; break;
        case 'f': out += '\f'; break;
        case 'v': out += '\v'; break;
        case '0': case '1': case '2': case '3': case '4': case '5': case '6': case '7': {      // \0, and legacy octal escapes \101
            unsigned code = (unsigned)(e - '0');
            for (int k = 0; k < 2 && pos < src.size() && src[pos] >= '0' && src[pos] <= '7' && code * 8 + (unsigned)(src[pos] - '0') <= 255; k++) code = code * 8 + (unsigned)(src[pos++] - '0');
            utf8(code);
            break;
        }
        case '\n': line++; break;                                   // line continuation
        case 'x': if (hex(2, v)) utf8(v); else out += 'x'; break;
        case 'u':
            if (pos < src.size() && src[pos] == '{') {
                size_t close = src.find('}', pos);
                unsigned cp = 0;
                bool ok = close != std::string::npos && close > pos + 1;
                for (size_t i = pos + 1; ok && i < close; i++) { if (!isxdigit((unsigned char)src[i])) ok = false; else cp = cp * 16 + (unsigned)(isdigit((unsigned char)src[i]) ? src[i] - '0' : tolower((unsigned char)src[i]) - 'a' + 10); }
                if (ok) { utf8(cp); pos = close + 1; } else out += 'u';
            } else if (hex(4, v)) utf8(v);
            else out += 'u';
            break;
        default: out += e;
    }
}

struct Lexer {
    const std::string& src;
    std::string name;
    size_t pos = 0;
    int line = 1;
    std::vector<Token> toks;

    Lexer(const std::string& s, const std::string& n) : src(s), name(n) {}

    [[noreturn]] void fail(const std::string& msg) {
        throw JsError(name + ":" + std::to_string(line) + ": SyntaxError: " + msg);
    }

    void run() {
        static const char* puncts[] = {
            ">>>=", "===", "!==", "**=", "<<=", ">>=", ">>>", "...", "=>", "==", "!=", "<=", ">=",
            "&&", "||", "??", "++", "--", "+=", "-=", "*=", "/=", "%=", "&=", "|=", "^=", "**",
            "<<", ">>", "{", "}", "(", ")", "[", "]", ";", ",", "<", ">", "+", "-", "*", "/", "%",
            "&", "|", "^", "!", "~", "?", ":", "=", "."};
        bool nl = false;
        while (true) {
            // whitespace and comments
            while (pos < src.size()) {
                char c = src[pos];
                if (c == '\n') { line++; nl = true; pos++; }
                else if (c == ' ' || c == '\t' || c == '\r' || c == '\f' || c == '\v') pos++;
                else if ((unsigned char)c == 0xEF && pos + 2 < src.size() && (unsigned char)src[pos + 1] == 0xBB) pos += 3;  // BOM
                else if (c == '/' && pos + 1 < src.size() && src[pos + 1] == '/') {
                    while (pos < src.size() && src[pos] != '\n') pos++;
                } else if (c == '/' && pos + 1 < src.size() && src[pos + 1] == '*') {
                    pos += 2;
                    while (pos + 1 < src.size() && !(src[pos] == '*' && src[pos + 1] == '/')) {
                        if (src[pos] == '\n') { line++; nl = true; }
                        pos++;
                    }
                    if (pos + 1 >= src.size()) fail("unterminated comment");
                    pos += 2;
                } else break;
            }
            Token t;
            t.line = line;
            t.nl_before = nl;
            nl = false;
            if (pos >= src.size()) { t.type = Tok::End; toks.push_back(t); return; }
            char c = src[pos];
            if (isdigit((unsigned char)c) || (c == '.' && pos + 1 < src.size() && isdigit((unsigned char)src[pos + 1]))) {
                const char kind = c == '0' && pos + 1 < src.size() ? (char)tolower((unsigned char)src[pos + 1]) : 0;
                if (kind == 'x' || kind == 'o' || kind == 'b') {              // 0x1f  0o17  0b101, with numeric separators
                    const int radix = kind == 'x' ? 16 : (kind == 'o' ? 8 : 2);
                    pos += 2;
                    double v = 0;
                    bool any = false;
                    for (; pos < src.size(); pos++) {
                        const char d = src[pos];
                        if (d == '_' && any) continue;
                        const int digit = isdigit((unsigned char)d) ? d - '0' : (isalpha((unsigned char)d) ? tolower((unsigned char)d) - 'a' + 10 : 99);
                        if (digit >= radix) break;
                        v = v * radix + digit;
                        any = true;
                    }
                    if (!any) fail("missing digits after the numeric prefix");
                    t.num = v;
                } else if (c == '0' && pos + 1 < src.size() && isdigit((unsigned char)src[pos + 1]) && [&] { size_t q = pos; while (q < src.size() && isdigit((unsigned char)src[q])) { if (src[q] > '7') return false; q++; } return q >= src.size() || (src[q] != '.' && src[q] != 'e' && src[q] != 'E'); }()) {
                    double v = 0;                                    // 010: a legacy octal literal (sloppy mode, as the reference's engine reads scripts)
                    for (; pos < src.size() && isdigit((unsigned char)src[pos]); pos++) v = v * 8 + (src[pos] - '0');
                    t.num = v;
                } else {
                    // decimal: the text with its numeric separators removed (1_000.5e1_0)
                    std::string text;
                    size_t q = pos;
                    bool seen_e = false, seen_dot = false;
                    for (; q < src.size(); q++) {
                        const char d = src[q];
                        if (isdigit((unsigned char)d)) text += d;
                        else if (d == '_' && !text.empty() && isdigit((unsigned char)text.back()) && q + 1 < src.size() && isdigit((unsigned char)src[q + 1])) continue;
                        else if (d == '.' && !seen_dot && !seen_e) { seen_dot = true; text += d; }
                        else if ((d == 'e' || d == 'E') && !seen_e && !text.empty() && q + 1 < src.size() &&
                                 (isdigit((unsigned char)src[q + 1]) || ((src[q + 1] == '+' || src[q + 1] == '-') && q + 2 < src.size() && isdigit((unsigned char)src[q + 2])))) {
                            seen_e = true;
                            text += d;
                            if (src[q + 1] == '+' || src[q + 1] == '-') text += src[++q];
                        } else break;
                    }
                    t.num = strtod(text.c_str(), nullptr);
                    pos = q;
                    if (pos < src.size() && src[pos] == 'n') pos++;  // BigInt suffix: treat as number
                }
                t.type = Tok::Num;
            } else if (isalpha((unsigned char)c) || c == '_' || c == '$') {
                size_t s = pos;
                while (pos < src.size() && (isalnum((unsigned char)src[pos]) || src[pos] == '_' || src[pos] == '$')) pos++;
                t.text = src.substr(s, pos - s);
                t.type = kKeywords.count(t.text) ? Tok::Keyword : Tok::Ident;
            } else if (c == '`') {
                // template literal: `a${x}b` is emitted as the tokens ("a" + (x) + "b"), the pieces lexed recursively
                if (!toks.empty() && !toks.back().nl_before && (toks.back().type == Tok::Ident || (toks.back().type == Tok::Punct && (toks.back().text == ")" || toks.back().text == "]"))))
                    fail("tagged template literals are not supported by the metric script interpreter");
                pos++;
                auto push_punct = [&](const char* s) { Token k; k.type = Tok::Punct; k.text = s; k.line = line; toks.push_back(k); };
                t.type = Tok::Punct; t.text = "("; toks.push_back(t);
                std::string piece;
                bool first = true;
                auto flush = [&]() {
                    if (!first) push_punct("+");
                    Token k; k.type = Tok::Str; k.text = piece; k.line = line; toks.push_back(k);
                    piece.clear();
                    first = false;
                };
                while (true) {
                    if (pos >= src.size()) fail("unterminated template literal");
                    char ch = src[pos++];
                    if (ch == '`') break;
                    if (ch == '\\' && pos < src.size()) {
                        read_escape(src, pos, piece, line);
                    } else if (ch == '$' && pos < src.size() && src[pos] == '{') {
                        pos++;
                        size_t start = pos;
                        int depth = 1;
                        while (pos < src.size() && depth > 0) {
                            if (src[pos] == '{') depth++;
                            else if (src[pos] == '}') depth--;
                            else if (src[pos] == '\n') line++;
                            if (depth > 0) pos++;
                        }
                        if (pos >= src.size()) fail("unterminated ${ in template literal");
                        std::string inner = src.substr(start, pos - start);
                        pos++;
                        flush();
                        push_punct("+");
                        { Token k; k.type = Tok::Ident; k.text = "$$template_string"; k.line = line; toks.push_back(k); }      // ToString of the value: toString() before valueOf()
                        push_punct("(");
                        Lexer sub(inner, name);
                        sub.line = line;
                        sub.run();
                        for (auto& k : sub.toks) if (k.type != Tok::End) toks.push_back(k);
                        push_punct(")");
                    } else {
                        if (ch == '\n') line++;
                        piece += ch;
                    }
                }
                flush();
                push_punct(")");
                continue;
            } else if (c == '"' || c == '\'') {
                char q = c;
                pos++;
                std::string out;
                while (pos < src.size() && src[pos] != q) {
                    char ch = src[pos++];
                    if (ch == '\n') { if (q != '`') fail("unterminated string"); line++; }
                    if (ch == '\\' && pos < src.size()) read_escape(src, pos, out, line);
                    else out += ch;
                }
                if (pos >= src.size()) fail("unterminated string");
                pos++;
                t.type = Tok::Str;
                t.text = out;
            } else {
                bool found = false;
                for (const char* p : puncts) {
                    size_t n = strlen(p);
                    if (src.compare(pos, n, p) == 0) {
                        t.type = Tok::Punct;
                        t.text = p;
                        pos += n;
                        found = true;
                        break;
                    }
                }
                if (!found) fail(std::string("unexpected character '") + c + "'");
            }
            toks.push_back(t);
        }
    }
};

// ------------------------------------------------------------------------------------------
// AST
// ------------------------------------------------------------------------------------------
enum class K {
    Num, Str, Ident, ArrayLit, ObjectLit, Func, Member, Index, Call, New, Unary, Update, Binary,
    Logical, Assign, Cond, Comma, This, Literal,
    VarDecl, FuncDecl, Return, If, For, ForIn, ForOf, While, DoWhile, Break, Continue, Block,
    ExprStmt, Empty, Throw, Try,
    Spread,      // ...expr in a call, an array literal or an object literal
    Switch,      // kids[0] discriminant, then Case nodes
    Case,        // kids[0] test (Empty for `default`), then the statements
    OptChain,    // kids[0] = a member / index / call chain with at least one `?.` link (flag on the link): undefined once a link's base is null or undefined
    Class,       // s = name, kids[0] = parent expression or Empty, members in `members`
    Labeled,     // s = label, kids[0] = the statement
    ObjectRest   // kids[0] = source object, s = excluded keys joined by '\x1f': the `...rest` of an object pattern
};

struct Node;
using NodeP = std::unique_ptr<Node>;

struct FuncDef {
    std::string name;
    std::vector<std::string> params;
    NodeP body;             // Block, or an expression for concise arrows
    bool is_arrow = false;
    bool expr_body = false;
    std::vector<NodeP> defaults;          // parallel to params (null: none): `function f(a, b = 2)`
    std::string rest;                     // `...rest`
    std::string source;                   // name of the script that defines the function (error messages of later calls)
    std::vector<NodeP> prologue;          // destructuring parameters: declarations run before the body
    bool strict = false;                  // class code, or under a "use strict" directive: `this` of a plain call stays undefined
    std::vector<std::string> hoisted_vars;
    std::vector<const Node*> hoisted_funcs;
};

struct Node {
    K k;
    int line = 0;
    std::string s;          // identifier / operator / string / decl kind
    double num = 0;
    bool flag = false;      // Update: prefix; Literal: bool value
    std::vector<NodeP> kids;
    std::unique_ptr<FuncDef> fn;
    struct ClassMember { std::string name; std::unique_ptr<FuncDef> fn; bool is_static = false; char accessor = 0; NodeP init; bool is_field = false; };   // accessor: 'g' getter, 's' setter; field: `init` (may be null)
    std::vector<ClassMember> members;       // Class
    Node(K kind, int ln) : k(kind), line(ln) {}
};

struct Program {
    std::string name;
    std::vector<NodeP> body;
    FuncDef top;  // holds hoisting info for the top level
};

static void collect_hoisted(const Node* n, FuncDef& into, bool top) {
    if (!n) return;
    switch (n->k) {
        case K::Func: return;  // do not descend into nested functions
        case K::FuncDecl:
            into.hoisted_funcs.push_back(n);
            return;
        case K::VarDecl:
            if (n->s == "var")
                for (auto& d : n->kids) into.hoisted_vars.push_back(d->s);
            for (auto& d : n->kids) for (auto& kk : d->kids) collect_hoisted(kk.get(), into, false);
            return;
        case K::ForIn: case K::ForOf:
            if (n->flag && n->num == 0) into.hoisted_vars.push_back(n->s);  // `for (var x in ..)`
            break;
        default: break;
    }
    for (auto& kk : n->kids) collect_hoisted(kk.get(), into, false);
    (void)top;
}

// ------------------------------------------------------------------------------------------
// Parser
// ------------------------------------------------------------------------------------------
struct Parser {
    std::vector<Token>& t;
    std::string name;
    size_t p = 0;
    Parser(std::vector<Token>& toks, const std::string& n) : t(toks), name(n) {}

    [[noreturn]] void fail(const std::string& msg) {
        throw JsError(name + ":" + std::to_string(t[p].line) + ": SyntaxError: " + msg +
                      (t[p].type == Tok::End ? " (at end of input)" : " (near '" + t[p].text + "')"));
    }
    const Token& cur() { return t[p]; }
    bool is_punct(const char* s) { return t[p].type == Tok::Punct && t[p].text == s; }
    bool is_kw(const char* s) { return t[p].type == Tok::Keyword && t[p].text == s; }
    bool eat_punct(const char* s) { if (is_punct(s)) { p++; return true; } return false; }
    bool eat_kw(const char* s) { if (is_kw(s)) { p++; return true; } return false; }
    void expect_punct(const char* s) { if (!eat_punct(s)) fail(std::string("expected '") + s + "'"); }
    NodeP mk(K k) { return std::make_unique<Node>(k, t[p].line); }

    void consume_semicolon() {
        if (eat_punct(";")) return;
        if (is_punct("}") || t[p].type == Tok::End || t[p].nl_before) return;  // ASI
        fail("expected ';'");
    }

    std::string ident_name() {
        if (t[p].type == Tok::Ident) return t[p++].text;
        // contextual keywords usable as identifiers
        if (t[p].type == Tok::Keyword && (t[p].text == "of" || t[p].text == "undefined")) return t[p++].text;
        fail("expected identifier");
    }
    std::string property_name() {
        if (t[p].type == Tok::Ident || t[p].type == Tok::Keyword) return t[p++].text;
        if (t[p].type == Tok::Str) return t[p++].text;
        if (t[p].type == Tok::Num) {
            char buf[32];
            double d = t[p++].num;
            if (d == std::floor(d)) snprintf(buf, sizeof buf, "%lld", (long long)d);
            else snprintf(buf, sizeof buf, "%.17g", d);
            return buf;
        }
        fail("expected property name");
    }

    std::unique_ptr<Program> parse_program() {
        auto prog = std::make_unique<Program>();
        prog->name = name;
        while (t[p].type != Tok::End) prog->body.push_back(statement());
        if (starts_with_use_strict(prog->body)) for (auto& s : prog->body) mark_strict(s.get());      // "use strict" at the top of the script
        for (auto& s : prog->body) collect_hoisted(s.get(), prog->top, true);
        return prog;
    }

    int tmp_counter = 0;
    std::string fresh(const char* stem) { return std::string("$") + stem + std::to_string(tmp_counter++); }
    NodeP ident_node(const std::string& nm) { auto n = mk(K::Ident); n->s = nm; return n; }

    // Binding pattern -> plain declarations.  `[a, b = 1, ...r]` and `{x, y: z = 3}` (nested patterns allowed) read from the
    // variable `src`; every bound name becomes an Ident declaration with its initialiser appended to `decls`.
    void bind_pattern(std::vector<NodeP>& decls, const std::string& src) {
        auto with_default = [&](std::function<NodeP()> access) -> NodeP {
            if (!eat_punct("=")) return access();
            auto test = mk(K::Binary);
            test->s = "===";
            test->kids.push_back(access());
            auto undef = mk(K::Literal);
            undef->s = "undefined";
            test->kids.push_back(std::move(undef));
            auto c = mk(K::Cond);
            c->kids.push_back(std::move(test));
            c->kids.push_back(assignment());
            c->kids.push_back(access());
            return c;
        };
        auto bind_target = [&](std::function<NodeP()> access) {
            if (is_punct("[") || is_punct("{")) {
                // nested pattern: a temporary holds the element (its own default applies to the whole sub-pattern)
                std::string tmp = fresh("d");
                std::vector<NodeP> inner;
                bind_pattern(inner, tmp);
                auto d = ident_node(tmp);
                d->kids.push_back(with_default(access));
                decls.push_back(std::move(d));
                for (auto& x : inner) decls.push_back(std::move(x));
                return;
            }
            auto d = ident_node(ident_name());
            d->kids.push_back(with_default(access));
            decls.push_back(std::move(d));
        };
        if (eat_punct("[")) {
            int index = 0;
            while (!is_punct("]")) {
                if (eat_punct(",")) { index++; continue; }          // hole
                if (eat_punct("...")) {
                    auto d = ident_node(ident_name());
                    auto call = mk(K::Call);
                    auto m = mk(K::Member);
                    m->s = "slice";
                    m->kids.push_back(ident_node(src));
                    call->kids.push_back(std::move(m));
                    auto from = mk(K::Num);
                    from->num = index;
                    call->kids.push_back(std::move(from));
                    d->kids.push_back(std::move(call));
                    decls.push_back(std::move(d));
                    break;
                }
                const int at = index++;
                bind_target([&, at]() { auto ix = mk(K::Index); ix->kids.push_back(ident_node(src)); auto k = mk(K::Num); k->num = at; ix->kids.push_back(std::move(k)); return NodeP(std::move(ix)); });
                if (!eat_punct(",")) break;
            }
            expect_punct("]");
        } else if (eat_punct("{")) {
            std::string taken;
            while (!is_punct("}")) {
                if (eat_punct("...")) {                    // {a, ...rest}: what is left of the source
                    auto d = ident_node(ident_name());
                    auto r = mk(K::ObjectRest);
                    r->s = taken;
                    r->kids.push_back(ident_node(src));
                    d->kids.push_back(std::move(r));
                    decls.push_back(std::move(d));
                    break;
                }
                const bool was_ident = t[p].type == Tok::Ident;
                std::string key = property_name();
                taken += key + '\x1f';
                auto access = [&, key]() { auto m = mk(K::Member); m->s = key; m->kids.push_back(ident_node(src)); return NodeP(std::move(m)); };
                if (eat_punct(":")) bind_target(access);
                else {
                    if (!was_ident) fail("expected ':' after a quoted or numeric property name in a pattern");
                    auto d = ident_node(key);
                    d->kids.push_back(with_default(access));
                    decls.push_back(std::move(d));
                }
                if (!eat_punct(",")) break;
            }
            expect_punct("}");
        } else fail("expected a binding pattern");
    }

    // formal parameters after '(' up to and including ')': names, defaults, a rest parameter, destructuring patterns
    void parameters(FuncDef& f) {
        while (!is_punct(")")) {
            if (eat_punct("...")) {
                f.rest = ident_name();
                break;
            }
            if (is_punct("[") || is_punct("{")) {
                std::string tmp = fresh("p");
                std::vector<NodeP> decls;
                bind_pattern(decls, tmp);
                f.params.push_back(tmp);
                f.defaults.push_back(eat_punct("=") ? assignment() : nullptr);
                auto d = mk(K::VarDecl);
                d->s = "let";
                d->kids = std::move(decls);
                f.prologue.push_back(std::move(d));
            } else {
                f.params.push_back(ident_name());
                f.defaults.push_back(eat_punct("=") ? assignment() : nullptr);
            }
            if (!eat_punct(",")) break;
        }
        expect_punct(")");
    }

    int strict_depth = 0;                 // > 0 inside class bodies and under "use strict"
    static bool starts_with_use_strict(const std::vector<NodeP>& body) {
        return !body.empty() && body[0]->k == K::ExprStmt && body[0]->kids[0]->k == K::Str && body[0]->kids[0]->s == "use strict";
    }
    std::unique_ptr<FuncDef> function_rest(const std::string& fname) {
        auto f = std::make_unique<FuncDef>();
        f->name = fname;
        f->source = name;
        if (is_punct("*")) fail("generator functions are not supported by the metric script interpreter");
        expect_punct("(");
        parameters(*f);
        f->body = block();
        // `let` / `const` at the top of the body must not redeclare a parameter (the reference's engine refuses the script)
        for (auto& st : f->body->kids) {
            if (st->k != K::VarDecl || st->s == "var") continue;
            for (auto& d : st->kids)
                if (std::find(f->params.begin(), f->params.end(), d->s) != f->params.end() || (!f->rest.empty() && d->s == f->rest))
                    throw JsError(name + ":" + std::to_string(st->line) + ": SyntaxError: invalid redefinition of parameter name '" + d->s + "'");
        }
        collect_hoisted(f->body.get(), *f, true);
        f->strict = strict_depth > 0 || starts_with_use_strict(f->body->kids);
        if (f->strict && strict_depth == 0) mark_strict(f->body.get());          // "use strict" in this function: the functions inside it too
        return f;
    }

    // functions nested in a body that turned out to be strict (its directive is only seen after the body is parsed)
    static void mark_strict(Node* n) {
        if (!n) return;
        if (n->fn) { n->fn->strict = true; mark_strict(n->fn->body.get()); for (auto& d : n->fn->defaults) mark_strict(d.get()); }
        for (auto& m : n->members) if (m.fn) { m.fn->strict = true; mark_strict(m.fn->body.get()); }
        for (auto& k : n->kids) mark_strict(k.get());
    }

    NodeP block() {
        auto b = mk(K::Block);
        expect_punct("{");
        while (!is_punct("}")) {
            if (t[p].type == Tok::End) fail("unterminated block");
            b->kids.push_back(statement());
        }
        expect_punct("}");
        check_lexical_redeclaration(b->kids);
        return b;
    }

    // two let / const / class declarations of one name in the same block (or one of them next to a function declaration of that name)
    void check_lexical_redeclaration(const std::vector<NodeP>& body) {
        std::set<std::string> lexical, functions;
        for (auto& st : body) {
            if (st->k == K::FuncDecl) { if (lexical.count(st->s)) throw JsError(name + ":" + std::to_string(st->line) + ": SyntaxError: invalid redefinition of lexical identifier '" + st->s + "'"); functions.insert(st->s); continue; }
            std::vector<std::string> names;
            if (st->k == K::VarDecl && st->s != "var") for (auto& d : st->kids) { if (d->s.compare(0, 1, "$") != 0 || d->s.find("$d") != 0) names.push_back(d->s); }
            else if (st->k == K::Class && !st->s.empty()) names.push_back(st->s);
            for (auto& nm : names)
                if (!lexical.insert(nm).second || functions.count(nm)) throw JsError(name + ":" + std::to_string(st->line) + ": SyntaxError: invalid redefinition of lexical identifier '" + nm + "'");
        }
    }

    NodeP var_decl(bool need_semicolon) {
        auto d = mk(K::VarDecl);
        d->s = t[p++].text;  // var/let/const
        do {
            if (is_punct("[") || is_punct("{")) {       // destructuring: const [a, b] = pair;  const {x, y} = point;
                std::string tmp = fresh("d");
                std::vector<NodeP> decls;
                bind_pattern(decls, tmp);
                if (!eat_punct("=")) fail("a destructuring declaration needs an initialiser");
                auto src = ident_node(tmp);
                src->kids.push_back(assignment());
                d->kids.push_back(std::move(src));
                for (auto& x : decls) d->kids.push_back(std::move(x));
                continue;
            }
            auto one = mk(K::Ident);
            one->s = ident_name();
            if (eat_punct("=")) one->kids.push_back(assignment());
            d->kids.push_back(std::move(one));
        } while (eat_punct(","));
        if (need_semicolon) consume_semicolon();
        return d;
    }

    NodeP statement() {
        if (is_punct("{")) return block();
        if (is_punct(";")) { auto e = mk(K::Empty); p++; return e; }
        if (is_kw("var") || is_kw("let") || is_kw("const")) return var_decl(true);
        if (is_kw("function")) {
            auto n = mk(K::FuncDecl);
            p++;
            if (is_punct("*")) fail("generator functions are not supported by the metric script interpreter");
            n->s = ident_name();
            n->fn = function_rest(n->s);
            return n;
        }
        if (is_kw("return")) {
            auto n = mk(K::Return);
            p++;
            if (!is_punct(";") && !is_punct("}") && t[p].type != Tok::End && !t[p].nl_before)
                n->kids.push_back(expression());
            consume_semicolon();
            return n;
        }
        if (is_kw("if")) {
            auto n = mk(K::If);
            p++;
            expect_punct("(");
            n->kids.push_back(expression());
            expect_punct(")");
            n->kids.push_back(statement());
            if (eat_kw("else")) n->kids.push_back(statement());
            return n;
        }
        if (is_kw("for")) {
            int ln = t[p].line;
            p++;
            expect_punct("(");
            // for-in / for-of detection
            size_t save = p;
            if (is_kw("var") || is_kw("let") || is_kw("const") || t[p].type == Tok::Ident) {
                bool decl = t[p].type == Tok::Keyword;
                std::string declkind = decl ? t[p].text : "";
                if (decl) p++;
                if (decl && (is_punct("[") || is_punct("{"))) {        // for (const [k, v] of pairs)
                    std::string tmp = fresh("it");
                    std::vector<NodeP> decls;
                    bind_pattern(decls, tmp);
                    if (!is_kw("of")) fail("a destructuring loop variable needs 'of'");
                    auto n = std::make_unique<Node>(K::ForOf, ln);
                    p++;
                    n->s = tmp;
                    n->flag = true;
                    n->num = 1;
                    n->kids.push_back(expression());
                    expect_punct(")");
                    auto body = mk(K::Block);
                    auto d = mk(K::VarDecl);
                    d->s = declkind == "var" ? "let" : declkind;
                    d->kids = std::move(decls);
                    body->kids.push_back(std::move(d));
                    body->kids.push_back(statement());
                    n->kids.push_back(std::move(body));
                    return n;
                }
                if (t[p].type == Tok::Ident) {
                    std::string vname = t[p].text;
                    p++;
                    if (is_kw("in") || is_kw("of")) {
                        auto n = std::make_unique<Node>(t[p].text == "in" ? K::ForIn : K::ForOf, ln);
                        p++;
                        n->s = vname;
                        n->flag = decl;
                        n->num = (declkind == "var") ? 0 : 1;
                        n->kids.push_back(expression());
                        expect_punct(")");
                        n->kids.push_back(statement());
                        return n;
                    }
                }
                p = save;
            }
            auto n = std::make_unique<Node>(K::For, ln);
            if (is_punct(";")) n->kids.push_back(mk(K::Empty));
            else if (is_kw("var") || is_kw("let") || is_kw("const")) n->kids.push_back(var_decl(false));
            else { auto e = mk(K::ExprStmt); e->kids.push_back(expression()); n->kids.push_back(std::move(e)); }
            expect_punct(";");
            if (is_punct(";")) n->kids.push_back(mk(K::Empty)); else n->kids.push_back(expression());
            expect_punct(";");
            if (is_punct(")")) n->kids.push_back(mk(K::Empty)); else n->kids.push_back(expression());
            expect_punct(")");
            n->kids.push_back(statement());
            return n;
        }
        if (is_kw("while")) {
            auto n = mk(K::While);
            p++;
            expect_punct("(");
            n->kids.push_back(expression());
            expect_punct(")");
            n->kids.push_back(statement());
            return n;
        }
        if (is_kw("do")) {
            auto n = mk(K::DoWhile);
            p++;
            n->kids.push_back(statement());
            if (!eat_kw("while")) fail("expected 'while'");
            expect_punct("(");
            n->kids.push_back(expression());
            expect_punct(")");
            eat_punct(";");
            return n;
        }
        if (is_kw("break") || is_kw("continue")) {
            auto n = mk(is_kw("break") ? K::Break : K::Continue);
            p++;
            if (t[p].type == Tok::Ident && !t[p].nl_before) n->s = t[p++].text;       // break label; / continue label;
            consume_semicolon();
            return n;
        }
        if (is_kw("throw")) {
            auto n = mk(K::Throw);
            p++;
            n->kids.push_back(expression());
            consume_semicolon();
            return n;
        }
        if (is_kw("try")) {
            auto n = mk(K::Try);
            p++;
            n->kids.push_back(block());
            if (eat_kw("catch")) {
                if (eat_punct("(")) { n->s = ident_name(); expect_punct(")"); }
                n->kids.push_back(block());
            } else n->kids.push_back(mk(K::Empty));
            if (eat_kw("finally")) n->kids.push_back(block());
            return n;
        }
        if (is_kw("switch")) {
            auto n = mk(K::Switch);
            p++;
            expect_punct("(");
            n->kids.push_back(expression());
            expect_punct(")");
            expect_punct("{");
            bool seen_default = false;
            while (!is_punct("}")) {
                auto c = mk(K::Case);
                if (eat_kw("default")) {
                    if (seen_default) fail("more than one default clause");
                    seen_default = true;
                    c->kids.push_back(mk(K::Empty));
                } else {
                    if (!eat_kw("case")) fail("expected 'case' or 'default'");
                    c->kids.push_back(expression());
                }
                expect_punct(":");
                while (!is_kw("case") && !is_kw("default") && !is_punct("}")) {
                    if (t[p].type == Tok::End) fail("unterminated switch");
                    c->kids.push_back(statement());
                }
                n->kids.push_back(std::move(c));
            }
            expect_punct("}");
            return n;
        }
        if (is_kw("class")) {
            auto c = class_rest();
            if (c->s.empty()) fail("a class declaration needs a name");
            return c;
        }
        if (t[p].type == Tok::Ident && t[p].text == "with" && t[p + 1].type == Tok::Punct && t[p + 1].text == "(")
            fail("'with' statements are not supported by the metric script interpreter");
        if (t[p].type == Tok::Ident && (t[p].text == "async" || t[p].text == "yield" || t[p].text == "await") && t[p + 1].type != Tok::Punct)
            fail("'" + t[p].text + "' is not supported by the metric script interpreter");
        if (t[p].type == Tok::Ident && t[p + 1].type == Tok::Punct && t[p + 1].text == ":") {
            auto n = mk(K::Labeled);
            n->s = t[p].text;
            p += 2;
            n->kids.push_back(statement());
            return n;
        }
        auto e = mk(K::ExprStmt);
        e->kids.push_back(expression());
        consume_semicolon();
        return e;
    }

    NodeP expression() {
        auto e = assignment();
        if (is_punct(",")) {
            auto c = mk(K::Comma);
            c->kids.push_back(std::move(e));
            while (eat_punct(",")) c->kids.push_back(assignment());
            return c;
        }
        return e;
    }

    // arrow function lookahead: `ident =>` or `( ... ) =>`
    bool looks_like_arrow() {
        if (t[p].type == Tok::Ident && t[p + 1].type == Tok::Punct && t[p + 1].text == "=>") return true;
        if (!is_punct("(")) return false;
        size_t q = p + 1;
        int depth = 1;
        while (q < t.size() && depth > 0) {
            if (t[q].type == Tok::End) return false;
            if (t[q].type == Tok::Punct) {
                if (t[q].text == "(") depth++;
                else if (t[q].text == ")") depth--;
            }
            q++;
        }
        return q < t.size() && t[q].type == Tok::Punct && t[q].text == "=>";
    }

    NodeP arrow_function() {
        auto n = mk(K::Func);
        auto f = std::make_unique<FuncDef>();
        f->is_arrow = true;
        f->source = name;
        if (t[p].type == Tok::Ident) { f->params.push_back(t[p++].text); f->defaults.push_back(nullptr); }
        else {
            expect_punct("(");
            parameters(*f);
        }
        expect_punct("=>");
        if (is_punct("{")) f->body = block();
        else { f->body = assignment(); f->expr_body = true; }
        collect_hoisted(f->body.get(), *f, true);
        n->fn = std::move(f);
        return n;
    }

    NodeP assignment() {
        if (looks_like_arrow()) return arrow_function();
        auto lhs = conditional();
        static const char* ops[] = {"=", "+=", "-=", "*=", "/=", "%=", "**=", "&=", "|=", "^=", "<<=", ">>=", ">>>="};
        for (const char* op : ops) {
            if (is_punct(op)) {
                const bool pattern = (lhs->k == K::ArrayLit || lhs->k == K::ObjectLit) && std::string(op) == "=";     // [a, b] = ...; ({a, b} = ...)
                if (lhs->k != K::Ident && lhs->k != K::Member && lhs->k != K::Index && !pattern) fail("invalid assignment target");
                auto n = mk(K::Assign);
                n->s = op;
                p++;
                n->kids.push_back(std::move(lhs));
                n->kids.push_back(assignment());
                return n;
            }
        }
        return lhs;
    }

    NodeP conditional() {
        auto c = binary(0);
        if (is_punct("?")) {
            auto n = mk(K::Cond);
            p++;
            n->kids.push_back(std::move(c));
            n->kids.push_back(assignment());
            expect_punct(":");
            n->kids.push_back(assignment());
            return n;
        }
        return c;
    }

    static int prec(const Token& tk) {
        if (tk.type == Tok::Keyword) {
            if (tk.text == "instanceof" || tk.text == "in") return 9;
            return -1;
        }
        if (tk.type != Tok::Punct) return -1;
        const std::string& s = tk.text;
        if (s == "??") return 1;
        if (s == "||") return 2;
        if (s == "&&") return 3;
        if (s == "|") return 4;
        if (s == "^") return 5;
        if (s == "&") return 6;
        if (s == "==" || s == "!=" || s == "===" || s == "!==") return 8;
        if (s == "<" || s == ">" || s == "<=" || s == ">=") return 9;
        if (s == "<<" || s == ">>" || s == ">>>") return 10;
        if (s == "+" || s == "-") return 11;
        if (s == "*" || s == "/" || s == "%") return 12;
        if (s == "**") return 13;
        return -1;
    }

    NodeP binary(int min_prec) {
        const bool bare_unary = (t[p].type == Tok::Punct && (t[p].text == "-" || t[p].text == "+" || t[p].text == "!" || t[p].text == "~")) || is_kw("typeof") || is_kw("void") || is_kw("delete");
        auto lhs = unary();
        if (bare_unary && is_punct("**")) fail("unparenthesized unary expression can't appear on the left-hand side of '**'");
        while (true) {
            int pr = prec(t[p]);
            if (pr < 0 || pr < min_prec) break;
            std::string op = t[p].text;
            int ln = t[p].line;
            p++;
            NodeP rhs = (op == "**") ? binary(pr) : binary(pr + 1);  // ** is right associative
            auto n = std::make_unique<Node>((op == "&&" || op == "||" || op == "??") ? K::Logical : K::Binary, ln);
            n->s = op;
            n->kids.push_back(std::move(lhs));
            n->kids.push_back(std::move(rhs));
            lhs = std::move(n);
        }
        return lhs;
    }

    NodeP unary() {
        if (t[p].type == Tok::Punct && (t[p].text == "-" || t[p].text == "+" || t[p].text == "!" || t[p].text == "~")) {
            auto n = mk(K::Unary);
            n->s = t[p++].text;
            n->kids.push_back(unary());
            return n;
        }
        if (is_kw("typeof") || is_kw("void") || is_kw("delete")) {
            auto n = mk(K::Unary);
            n->s = t[p++].text;
            n->kids.push_back(unary());
            return n;
        }
        if (is_punct("++") || is_punct("--")) {
            auto n = mk(K::Update);
            n->s = t[p++].text;
            n->flag = true;
            n->kids.push_back(unary());
            return n;
        }
        auto e = postfix();
        return e;
    }

    NodeP postfix() {
        auto e = call_member();
        if ((is_punct("++") || is_punct("--")) && !t[p].nl_before) {
            auto n = mk(K::Update);
            n->s = t[p++].text;
            n->flag = false;
            n->kids.push_back(std::move(e));
            return n;
        }
        return e;
    }

    NodeP maybe_spread() {
        if (is_punct("...")) {
            auto sp = mk(K::Spread);
            p++;
            sp->kids.push_back(assignment());
            return sp;
        }
        return assignment();
    }

    void arguments(Node& call) {
        expect_punct("(");
        while (!is_punct(")")) {
            call.kids.push_back(maybe_spread());
            if (!eat_punct(",")) break;
        }
        expect_punct(")");
    }

    // after the `class` keyword: [name] [extends expr] { constructor(..){..}  m(..){..}  static s(..){..}  get x(){..}  set x(v){..} }
    NodeP class_rest() {
        auto c = mk(K::Class);
        p++;
        if (t[p].type == Tok::Ident) c->s = t[p++].text;
        if (t[p].type == Tok::Ident && t[p].text == "extends") { p++; c->kids.push_back(call_member()); }
        else c->kids.push_back(mk(K::Empty));
        expect_punct("{");
        strict_depth++;                                   // class code is strict
        struct Leave { int& d; ~Leave() { d--; } } leave{strict_depth};
        while (!is_punct("}")) {
            if (eat_punct(";")) continue;
            Node::ClassMember m;
            if (t[p].type == Tok::Ident && t[p].text == "static" && !(t[p + 1].type == Tok::Punct && t[p + 1].text == "(")) { m.is_static = true; p++; }
            if (t[p].type == Tok::Ident && (t[p].text == "get" || t[p].text == "set") && !(t[p + 1].type == Tok::Punct && t[p + 1].text == "(")) { m.accessor = t[p].text[0]; p++; }
            if (is_punct("*") || is_punct("[") || is_punct("#")) fail("generator, computed and private class members are not supported by the metric script interpreter");
            m.name = property_name();
            if (!is_punct("(")) {                          // a field: `x = expr;`, `x;`, `static y = expr;`
                if (m.accessor) fail("expected '(' after an accessor name");
                m.is_field = true;
                if (eat_punct("=")) m.init = assignment();
                consume_semicolon();
                c->members.push_back(std::move(m));
                continue;
            }
            m.fn = function_rest(m.name);
            if (m.accessor && (m.fn->params.size() != (m.accessor == 'g' ? 0u : 1u) || !m.fn->rest.empty())) fail("invalid number of arguments for getter or setter");
            c->members.push_back(std::move(m));
        }
        expect_punct("}");
        lower_instance_fields(*c);
        return c;
    }

    // Instance fields become assignments `this.name = init;` in the constructor: at the top of a base class's constructor, right
    // behind the top-level `super(...)` statement of a derived class's (where JavaScript runs them); a class without a constructor
    // gets the default one.  (An initialiser therefore sees the constructor's parameters, which JavaScript's does not - a script
    // that shadows an outer name by a constructor parameter AND reads that name in a field initialiser would differ.)  Static
    // fields stay members: the evaluator sets them when the class is created.
    void lower_instance_fields(Node& c) {
        std::vector<NodeP> inits;
        for (auto it = c.members.begin(); it != c.members.end();) {
            if (!it->is_field || it->is_static) { ++it; continue; }
            auto target = mk(K::Member);
            target->s = it->name;
            target->kids.push_back(mk(K::This));
            auto asg = mk(K::Assign);
            asg->s = "=";
            asg->kids.push_back(std::move(target));
            if (it->init) asg->kids.push_back(std::move(it->init));
            else { auto u = mk(K::Literal); u->s = "undefined"; asg->kids.push_back(std::move(u)); }
            auto st = mk(K::ExprStmt);
            st->kids.push_back(std::move(asg));
            inits.push_back(std::move(st));
            it = c.members.erase(it);
        }
        if (inits.empty()) return;
        const bool derived = c.kids[0]->k != K::Empty;
        FuncDef* ctor = nullptr;
        for (auto& m : c.members)
            if (m.name == "constructor" && !m.is_static && !m.accessor && !m.is_field) ctor = m.fn.get();
        if (!ctor) {
            Node::ClassMember m;
            m.name = "constructor";
            m.fn = std::make_unique<FuncDef>();
            m.fn->name = "constructor";
            m.fn->source = name;
            m.fn->body = mk(K::Block);
            if (derived) {                                   // constructor(...args) { super(...args); }
                m.fn->rest = "args";
                auto call = mk(K::Call);
                auto callee = mk(K::Ident);
                callee->s = "super";
                call->kids.push_back(std::move(callee));
                auto sp = mk(K::Spread);
                auto a = mk(K::Ident);
                a->s = "args";
                sp->kids.push_back(std::move(a));
                call->kids.push_back(std::move(sp));
                auto st = mk(K::ExprStmt);
                st->kids.push_back(std::move(call));
                m.fn->body->kids.push_back(std::move(st));
            }
            ctor = m.fn.get();
            c.members.push_back(std::move(m));
        }
        auto& body = ctor->body->kids;
        size_t at = 0;
        if (derived) {
            at = body.size() + 1;
            for (size_t i = 0; i < body.size(); i++) {
                const Node* st = body[i].get();
                if (st->k == K::ExprStmt && st->kids[0]->k == K::Call && st->kids[0]->kids[0]->k == K::Ident && st->kids[0]->kids[0]->s == "super") { at = i + 1; break; }
            }
            if (at > body.size()) fail("class fields of a derived class need super(...) as a top-level statement of its constructor");
        }
        for (size_t i = 0; i < inits.size(); i++) body.insert(body.begin() + (long)(at + i), std::move(inits[i]));
    }

    NodeP call_member() {
        NodeP e;
        if (is_kw("new")) {
            if (t[p + 1].type == Tok::Punct && t[p + 1].text == ".") fail("new.target is not supported by the metric script interpreter");
            auto n = mk(K::New);
            p++;
            // callee without call parens
            NodeP callee = primary();
            while (true) {
                if (eat_punct(".")) {
                    auto m = mk(K::Member);
                    m->s = property_name();
                    m->kids.push_back(std::move(callee));
                    callee = std::move(m);
                } else break;
            }
            n->kids.push_back(std::move(callee));
            if (is_punct("(")) arguments(*n);
            e = std::move(n);
        } else e = primary();
        bool optional_chain = false;
        while (true) {
            if (is_punct("?") && t[p + 1].type == Tok::Punct && t[p + 1].text == ".") {      // a?.b   a?.[k]   a?.(args)
                p += 2;
                optional_chain = true;
                if (is_punct("[")) {
                    auto m = mk(K::Index);
                    p++;
                    m->flag = true;
                    m->kids.push_back(std::move(e));
                    m->kids.push_back(expression());
                    expect_punct("]");
                    e = std::move(m);
                } else if (is_punct("(")) {
                    auto c = mk(K::Call);
                    c->flag = true;
                    c->kids.push_back(std::move(e));
                    arguments(*c);
                    e = std::move(c);
                } else {
                    auto m = mk(K::Member);
                    m->flag = true;
                    m->s = property_name();
                    m->kids.push_back(std::move(e));
                    e = std::move(m);
                }
                continue;
            }
            if (is_punct(".")) {
                auto m = mk(K::Member);
                p++;
                m->s = property_name();
                m->kids.push_back(std::move(e));
                e = std::move(m);
            } else if (is_punct("[")) {
                auto m = mk(K::Index);
                p++;
                m->kids.push_back(std::move(e));
                m->kids.push_back(expression());
                expect_punct("]");
                e = std::move(m);
            } else if (is_punct("(")) {
                auto c = mk(K::Call);
                c->kids.push_back(std::move(e));
                arguments(*c);
                e = std::move(c);
            } else break;
        }
        if (optional_chain) {
            auto w = mk(K::OptChain);
            w->kids.push_back(std::move(e));
            e = std::move(w);
        }
        return e;
    }

    NodeP primary() {
        const Token& tk = t[p];
        if (tk.type == Tok::Num) { auto n = mk(K::Num); n->num = tk.num; p++; return n; }
        if (tk.type == Tok::Str) { auto n = mk(K::Str); n->s = tk.text; p++; return n; }
        if (tk.type == Tok::Ident) { auto n = mk(K::Ident); n->s = tk.text; p++; return n; }
        if (tk.type == Tok::Keyword) {
            if (tk.text == "this") { auto n = mk(K::This); p++; return n; }
            if (tk.text == "true" || tk.text == "false") { auto n = mk(K::Literal); n->s = "bool"; n->flag = tk.text == "true"; p++; return n; }
            if (tk.text == "null") { auto n = mk(K::Literal); n->s = "null"; p++; return n; }
            if (tk.text == "undefined") { auto n = mk(K::Literal); n->s = "undefined"; p++; return n; }
            if (tk.text == "of") { auto n = mk(K::Ident); n->s = tk.text; p++; return n; }
            if (tk.text == "class") return class_rest();
            if (tk.text == "function") {
                auto n = mk(K::Func);
                p++;
                if (is_punct("*")) fail("generator functions are not supported by the metric script interpreter");
                std::string fname;
                if (t[p].type == Tok::Ident) fname = t[p++].text;
                n->fn = function_rest(fname);
                return n;
            }
        }
        if (is_punct("(")) {
            p++;
            auto e = expression();
            expect_punct(")");
            return e;
        }
        if (is_punct("[")) {
            auto a = mk(K::ArrayLit);
            p++;
            while (!is_punct("]")) {
                if (is_punct(",")) { a->kids.push_back(mk(K::Literal)); a->kids.back()->s = "hole"; p++; continue; }
                a->kids.push_back(maybe_spread());
                if (!eat_punct(",")) break;
            }
            expect_punct("]");
            return a;
        }
        if (is_punct("{")) {
            auto o = mk(K::ObjectLit);
            p++;
            while (!is_punct("}")) {
                if (is_punct("...")) {                    // {...base, x: 1}
                    auto key = mk(K::Str);
                    key->s = "...";
                    p++;
                    o->kids.push_back(std::move(key));
                    o->kids.push_back(assignment());
                    if (!eat_punct(",")) break;
                    continue;
                }
                if (eat_punct("[")) {                     // {[expr]: value} and {[expr](..) {..}}: the key node is the expression itself
                    NodeP key = assignment();
                    expect_punct("]");
                    if (key->k == K::Str) { auto wrap = mk(K::Comma); wrap->kids.push_back(std::move(key)); key = std::move(wrap); }   // a Str node is a literal key: keep a computed one apart
                    if (is_punct("(")) {
                        auto fnode = mk(K::Func);
                        fnode->fn = function_rest("");
                        o->kids.push_back(std::move(key));
                        o->kids.push_back(std::move(fnode));
                    } else {
                        expect_punct(":");
                        o->kids.push_back(std::move(key));
                        o->kids.push_back(assignment());
                    }
                    if (!eat_punct(",")) break;
                    continue;
                }
                if (t[p].type == Tok::Ident && (t[p].text == "get" || t[p].text == "set") &&
                    !(t[p + 1].type == Tok::Punct && (t[p + 1].text == "(" || t[p + 1].text == ":" || t[p + 1].text == "," || t[p + 1].text == "}" || t[p + 1].text == "="))) {
                    // {get v() {...}, set v(x) {...}}: stored like a class's accessors, under "$get$v" / "$set$v"
                    const char which = t[p].text[0];
                    p++;
                    auto key = mk(K::Str);
                    key->s = std::string(which == 'g' ? "$get$" : "$set$") + property_name();
                    auto fnode = mk(K::Func);
                    fnode->fn = function_rest(key->s.substr(5));
                    if (fnode->fn->params.size() != (which == 'g' ? 0u : 1u) || !fnode->fn->rest.empty()) fail("invalid number of arguments for getter or setter");
                    o->kids.push_back(std::move(key));
                    o->kids.push_back(std::move(fnode));
                    if (!eat_punct(",")) break;
                    continue;
                }
                auto key = mk(K::Str);
                bool was_ident = t[p].type == Tok::Ident;
                key->s = property_name();
                if (was_ident && is_punct("=")) {          // {x = 1}: only meaningful as an assignment pattern ({x = 1} = obj)
                    auto id = mk(K::Ident);
                    id->s = key->s;
                    auto dflt = mk(K::Assign);
                    dflt->s = "=";
                    dflt->flag = true;                     // a shorthand default, not an assignment expression
                    p++;
                    dflt->kids.push_back(std::move(id));
                    dflt->kids.push_back(assignment());
                    o->kids.push_back(std::move(key));
                    o->kids.push_back(std::move(dflt));
                    if (!eat_punct(",")) break;
                    continue;
                }
                if (is_punct("(")) {  // method shorthand
                    auto fnode = mk(K::Func);
                    fnode->fn = function_rest(key->s);
                    o->kids.push_back(std::move(key));
                    o->kids.push_back(std::move(fnode));
                } else if (was_ident && (is_punct(",") || is_punct("}"))) {  // {x}
                    auto id = mk(K::Ident);
                    id->s = key->s;
                    o->kids.push_back(std::move(key));
                    o->kids.push_back(std::move(id));
                } else {
                    expect_punct(":");
                    o->kids.push_back(std::move(key));
                    o->kids.push_back(assignment());
                }
                if (!eat_punct(",")) break;
            }
            expect_punct("}");
            return o;
        }
        if (is_punct("/") || is_punct("/=")) fail("regular expression literals are not supported by the metric script interpreter");
        fail("unexpected token");
    }
};

// ------------------------------------------------------------------------------------------
// Evaluator
// ------------------------------------------------------------------------------------------
enum class Comp { Normal, Break, Continue, Return };

struct Thrown {
    Value v;
};

struct Evaluator {
    Interp& in;
    std::string name;
    Value ret;         // value carried by Return
    Value completion;  // completion value of the last expression statement
    Evaluator(Interp& i, const std::string& n) : in(i), name(n) {}

    [[noreturn]] void fail(const Node* n, const std::string& msg) {
        throw JsError(name + ":" + std::to_string(n ? n->line : 0) + ": " + msg);
    }
    void tick(const Node* n) {
        if (++in.steps > 50000000L) fail(n, "script exceeded the execution step limit");
    }

    // ---- environments ----
    static EnvP new_env(EnvP parent, bool fn_scope) {
        auto e = std::make_shared<Env>();
        e->parent = std::move(parent);
        e->function_scope = fn_scope;
        if (e->parent) e->this_val = e->parent->this_val;
        return e;
    }
    Env* lookup_env(Env* e, const std::string& nm) {
        for (; e; e = e->parent.get())
            if (e->vars.count(nm)) return e;
        return nullptr;
    }
    Value get_var(const Node* n, EnvP& env, const std::string& nm) {
        Env* e = lookup_env(env.get(), nm);
        if (e) return e->vars[nm];
        auto g = in.global_getters.find(nm);
        if (g != in.global_getters.end()) return g->second(in);
        fail(n, "ReferenceError: '" + nm + "' is not defined");
    }
    void set_var(const Node* n, EnvP& env, const std::string& nm, const Value& v) {
        Env* e = lookup_env(env.get(), nm);
        if (!e) {
            if (in.global_getters.count(nm)) return;  // e.g. `$cfg = ...`: setter is a no-op in the reference
            in.globals->vars[nm] = v;                 // sloppy-mode implicit global
            return;
        }
        if (e->is_const.count(nm) && e->is_const[nm]) fail(n, "TypeError: assignment to constant variable '" + nm + "'");
        e->vars[nm] = v;
    }
    void hoist(const FuncDef& f, EnvP& env) {
        for (auto& v : f.hoisted_vars)
            if (!env->vars.count(v)) env->vars[v] = Value();
        for (const Node* fd : f.hoisted_funcs) env->vars[fd->s] = make_closure(fd->fn.get(), env);
    }
    Value make_closure(const FuncDef* def, EnvP& env) {
        auto o = std::make_shared<Object>();
        o->kind = Object::Function;
        o->def = def;
        o->env = env;
        if (def->is_arrow) { o->is_arrow = true; o->bound_this = env->this_val; }
        return Value(o);
    }

    // ---- property access ----
    static bool array_index(const std::string& key, size_t& idx) {
        if (key.empty() || key.size() > 10) return false;
        for (char c : key) if (!isdigit((unsigned char)c)) return false;
        idx = (size_t)strtoull(key.c_str(), nullptr, 10);
        return true;
    }
    std::string key_of(const Value& v) {
        if (v.type == Value::Number) {
            char buf[32];
            if (v.num == std::floor(v.num) && std::fabs(v.num) < 1e15) snprintf(buf, sizeof buf, "%lld", (long long)v.num);
            else snprintf(buf, sizeof buf, "%.17g", v.num);
            return buf;
        }
        return in.to_string(v);
    }
    Value get_prop(const Node* n, const Value& objv, const std::string& key);
    void set_prop(const Node* n, const Value& objv, const std::string& key, const Value& v);
    Value array_method(const Node* n, const ObjP& a, const std::string& key);
    Value string_method(const std::string& str, const std::string& key);
    // valueOf(), then toString(), of a plain object or class instance that defines them; anything else comes back unchanged
    Value to_primitive(const Node* n, const Value& v) {
        if (v.type != Value::Obj || v.obj->kind != Object::Plain) return v;
        for (const char* name : {"valueOf", "toString"}) {
            const Object* h = holder_of(v.obj.get(), name);           // its own or its class's
            if (!h) continue;
            const Value& fn = h->props.find(name)->second;
            if (fn.type != Value::Obj || (fn.obj->kind != Object::Function && fn.obj->kind != Object::Native)) continue;
            std::vector<Value> none;
            Value r = call_function(n, fn, v, none);
            if (r.type != Value::Obj) return r;
        }
        return v;
    }
    Value number_method(double num, const std::string& key);

    // ---- operators ----
    bool is_host(const Value& v) { return v.type == Value::Obj && v.obj->kind == Object::Host; }
    Value binary_op(const Node* n, const std::string& op, const Value& a, const Value& b);
    bool loose_equal(const Value& a, const Value& b);
    bool strict_equal(const Value& a, const Value& b);

    // break / continue with a label: `jump_label` names the target while the completion travels outwards; a loop learns the labels it
    // carries from the Labeled node that runs it
    bool chain_short = false;        // an optional chain met a null / undefined base: the links above it yield undefined (K::OptChain resets it)
    std::string jump_label;
    std::vector<std::string> loop_labels;
    std::vector<std::string> take_loop_labels() { std::vector<std::string> l; l.swap(loop_labels); return l; }
    // what a loop does with its body's completion: true = stop iterating (the caller breaks out normally when it was this loop's own
    // break, otherwise hands the completion on); false = next iteration (normal completion, or a continue meant for this loop)
    bool leaves_loop(Comp c, const std::vector<std::string>& mine) {
        const bool own = jump_label.empty() || std::find(mine.begin(), mine.end(), jump_label) != mine.end();
        if (c == Comp::Normal) return false;
        if (c == Comp::Return) return true;
        if (own) {
            jump_label.clear();
            return c == Comp::Break;          // own continue: keep going
        }
        return true;                          // a break / continue for an outer statement: out of this loop, label kept
    }
    Value eval(const Node* n, EnvP& env);
    Comp exec(const Node* n, EnvP& env);
    Comp exec_block(const std::vector<NodeP>& body, EnvP& env);
    Value call_function(const Node* n, const Value& fn, const Value& self, std::vector<Value>& args);
    void assign_to(const Node* target, EnvP& env, const Value& v);
    // assignment pattern: an array or object literal on the left of '=' (elements: targets, nested patterns, `target = default`,
    // holes, a trailing `...rest`)
    void destructure_assign(const Node* pat, EnvP& env, const Value& v) {
        auto put = [&](const Node* el, Value item) {
            const Node* target = el;
            if (el->k == K::Assign && el->s == "=") {                    // default value
                target = el->kids[0].get();
                if (item.is_undefined()) item = eval(el->kids[1].get(), env);
            }
            if (target->k == K::ArrayLit || target->k == K::ObjectLit) destructure_assign(target, env, item);
            else if (target->k == K::Ident || target->k == K::Member || target->k == K::Index) assign_to(target, env, item);
            else fail(target, "SyntaxError: invalid destructuring assignment target");
        };
        if (v.type == Value::Undefined || v.type == Value::Null) fail(pat, "TypeError: cannot destructure 'undefined' or 'null'");
        if (pat->k == K::ArrayLit) {
            std::vector<Value> items;
            if (v.type == Value::Obj && v.obj->kind == Object::Array) items = v.obj->arr;
            else if (v.type == Value::String) for (char ch : v.str) items.push_back(Value(std::string(1, ch)));
            else fail(pat, "TypeError: array destructuring needs an array");
            for (size_t i = 0; i < pat->kids.size(); i++) {
                const Node* el = pat->kids[i].get();
                if (el->k == K::Literal && el->s == "hole") continue;
                if (el->k == K::Spread) {
                    std::vector<Value> rest(items.begin() + (long)std::min(i, items.size()), items.end());
                    put(el->kids[0].get(), Value(Interp::make_array(rest)));
                    break;
                }
                put(el, i < items.size() ? items[i] : Value());
            }
            return;
        }
        std::string taken;
        for (size_t i = 0; i + 1 < pat->kids.size(); i += 2) {
            const Node* key = pat->kids[i].get();
            const Node* el = pat->kids[i + 1].get();
            if (key->k == K::Str && key->s == "..." && el->k != K::Str) {       // {...rest}
                auto o = Interp::make_object();
                if (v.type == Value::Obj)
                    for (auto& kv : v.obj->props)
                        if (taken.find('\x1f' + kv.first + '\x1f') == std::string::npos) o->props[kv.first] = kv.second;
                put(el, Value(o));
                break;
            }
            const std::string name = key->k == K::Str ? key->s : key_of(eval(key, env));
            taken += '\x1f' + name + '\x1f';
            put(el, get_prop(pat, v, name));
        }
    }

    // one argument / array element: `...expr` contributes every element of the array it evaluates to
    void push_arg(const Node* k, EnvP& env, std::vector<Value>& out) {
        if (k->k != K::Spread) { out.push_back(eval(k, env)); return; }
        Value v = eval(k->kids[0].get(), env);
        std::vector<Value> items;
        if (!Interp::iterate(v, items)) fail(k, "TypeError: spread needs an array, a string, a Map or a Set");
        out.insert(out.end(), items.begin(), items.end());
    }

    // F.prototype of a function: made on first use for a plain function (ES5 constructors: F.prototype.m = function ...), the class's
    // prototype object for a class
    Value prototype_of_function(const ObjP& f) {
        if (!f || f->kind != Object::Function) return Value();
        auto it = f->props.find("prototype");
        if (it != f->props.end()) return it->second;
        if (f->is_arrow) return Value();
        Value p(Interp::make_object());
        f->props["prototype"] = p;
        return p;
    }
    // the first object of o's prototype chain (o included) that has `key` as its own property; null if none
    static const Object* holder_of(const Object* o, const std::string& key) {
        for (int depth = 0; o && depth < 64; o = o->proto.get(), depth++)
            if (o->props.count(key)) return o;
        return nullptr;
    }

    // class C [extends B] { ... } -> a constructor object: props["prototype"] holds the methods (getters / setters under "$get$name" /
    // "$set$name") and chains to the parent's prototype, static members are properties of the constructor, which chains to the parent class
    Value make_class(const Node* n, EnvP& env) {
        Value parent;
        if (n->kids[0]->k != K::Empty) {
            parent = eval(n->kids[0].get(), env);
            const bool native_error = parent.type == Value::Obj && parent.obj->kind == Object::Native && [&] { for (auto& kv : in.error_types) if (kv.second == parent.obj) return true; return false; }();
            if (parent.type != Value::Obj || (parent.obj->kind != Object::Function && !native_error)) fail(n, "TypeError: class extends value is not a constructor");
        }
        EnvP cenv = new_env(env, false);        // `super` and the class's own name are visible to its methods
        auto proto = Interp::make_object();
        auto ctor = std::make_shared<Object>();
        ctor->kind = Object::Function;
        ctor->is_class = true;
        ctor->env = cenv;
        const FuncDef* ctor_def = nullptr;
        proto->class_proto = true;
        if (parent.type == Value::Obj) {
            ctor->klass = parent.obj;
            ctor->proto = parent.obj;                                                       // inherited statics
            Value pp = prototype_of_function(parent.obj);
            if (pp.type == Value::Obj) proto->proto = pp.obj;                               // inherited methods
            // `super`: callable (the parent constructor on the current `this`); its properties are the parent's prototype's
            auto sup = std::make_shared<Object>(*parent.obj);
            sup->is_class = false;
            sup->props.clear();
            sup->proto = pp.type == Value::Obj ? pp.obj : nullptr;
            cenv->vars["super"] = Value(sup);
        }
        for (auto& m : n->members) {
            if (m.is_field) continue;                     // static fields: below, once the class is complete (instance fields were lowered by the parser)
            if (m.name == "constructor" && !m.is_static && !m.accessor) { ctor_def = m.fn.get(); continue; }
            auto f = std::make_shared<Object>();
            f->kind = Object::Function;
            f->def = m.fn.get();
            f->env = cenv;
            std::string key = m.accessor == 'g' ? "$get$" + m.name : (m.accessor == 's' ? "$set$" + m.name : m.name);
            (m.is_static ? ctor->props : proto->props)[key] = Value(f);
        }
        static const FuncDef empty_ctor = [] { FuncDef d; d.body = std::make_unique<Node>(K::Block, 0); return d; }();
        if (!ctor_def && parent.type == Value::Obj) {
            // default derived constructor: constructor(...args) { super(...args); }
            ctor->def = parent.obj->def ? parent.obj->def : &empty_ctor;
            ctor->env = parent.obj->def ? parent.obj->env : cenv;
        } else ctor->def = ctor_def ? ctor_def : &empty_ctor;
        if (parent.type == Value::Obj && parent.obj->kind == Object::Native) proto->klass = parent.obj;      // class E extends Error: instances are Errors
        ctor->props["prototype"] = Value(proto);
        if (!n->s.empty() && !ctor->props.count("name")) ctor->props["name"] = Value(n->s);
        Value cv(ctor);
        if (!n->s.empty()) cenv->vars[n->s] = cv;
        for (auto& m : n->members) {                      // static x = expr;  (`this` is the class)
            if (!m.is_field) continue;
            EnvP fenv = new_env(cenv, false);
            fenv->this_val = cv;
            ctor->props[m.name] = m.init ? eval(m.init.get(), fenv) : Value();
        }
        return cv;
    }
};

Value Evaluator::array_method(const Node* n, const ObjP& a, const std::string& key) {
    std::weak_ptr<Object> wa = a;
    auto self = [wa]() { return wa.lock(); };
    if (key == "push") return Interp::make_native([self](Interp&, const Value&, std::vector<Value>& args) {
        auto s = self();
        for (auto& v : args) s->arr.push_back(v);
        return Value((double)s->arr.size());
    });
    if (key == "pop") return Interp::make_native([self](Interp&, const Value&, std::vector<Value>&) {
        auto s = self();
        if (s->arr.empty()) return Value();
        Value v = s->arr.back();
        s->arr.pop_back();
        return v;
    });
    if (key == "slice") return Interp::make_native([self](Interp&, const Value&, std::vector<Value>& args) {
        auto s = self();
        long len = (long)s->arr.size();
        long b = args.size() > 0 ? (long)Interp::to_number(args[0]) : 0;
        long e = args.size() > 1 && !args[1].is_undefined() ? (long)Interp::to_number(args[1]) : len;
        if (b < 0) b += len;
        if (e < 0) e += len;
        b = std::max(0L, std::min(b, len));
        e = std::max(b, std::min(e, len));
        return Value(Interp::make_array(std::vector<Value>(s->arr.begin() + b, s->arr.begin() + e)));
    });
    if (key == "concat") return Interp::make_native([self](Interp&, const Value&, std::vector<Value>& args) {
        auto s = self();
        std::vector<Value> out = s->arr;
        for (auto& v : args) {
            if (v.type == Value::Obj && v.obj->kind == Object::Array) out.insert(out.end(), v.obj->arr.begin(), v.obj->arr.end());
            else out.push_back(v);
        }
        return Value(Interp::make_array(out));
    });
    if (key == "fill") return Interp::make_native([self](Interp&, const Value&, std::vector<Value>& args) {
        auto s = self();
        for (auto& v : s->arr) v = args.empty() ? Value() : args[0];
        return Value(s);
    });
    if (key == "indexOf") return Interp::make_native([self, this](Interp&, const Value&, std::vector<Value>& args) {
        auto s = self();
        for (size_t i = 0; i < s->arr.size(); i++)
            if (!args.empty() && strict_equal(s->arr[i], args[0])) return Value((double)i);
        return Value(-1.0);
    });
    if (key == "join") return Interp::make_native([self](Interp& ip, const Value&, std::vector<Value>& args) {
        auto s = self();
        std::string sep = args.empty() ? "," : ip.to_string(args[0]);
        std::string out;
        for (size_t i = 0; i < s->arr.size(); i++) { if (i) out += sep; if (s->arr[i].type != Value::Undefined && s->arr[i].type != Value::Null) out += ip.to_string(s->arr[i]); }
        return Value(out);
    });
    if (key == "reverse") return Interp::make_native([self](Interp&, const Value&, std::vector<Value>&) {
        auto s = self();
        std::reverse(s->arr.begin(), s->arr.end());
        return Value(s);
    });
    if (key == "shift") return Interp::make_native([self](Interp&, const Value&, std::vector<Value>&) {
        auto s = self();
        if (s->arr.empty()) return Value();
        Value v = s->arr.front();
        s->arr.erase(s->arr.begin());
        return v;
    });
    if (key == "unshift") return Interp::make_native([self](Interp&, const Value&, std::vector<Value>& args) {
        auto s = self();
        s->arr.insert(s->arr.begin(), args.begin(), args.end());
        return Value((double)s->arr.size());
    });
    if (key == "splice") return Interp::make_native([self](Interp&, const Value&, std::vector<Value>& args) {
        auto s = self();
        const long len = (long)s->arr.size();
        long b = args.size() > 0 ? (long)Interp::to_number(args[0]) : 0;
        if (b < 0) b = std::max(0L, len + b);
        b = std::min(b, len);
        long cnt = args.size() > 1 ? (long)Interp::to_number(args[1]) : len - b;
        cnt = std::max(0L, std::min(cnt, len - b));
        if (args.empty()) cnt = 0;
        std::vector<Value> removed(s->arr.begin() + b, s->arr.begin() + b + cnt);
        s->arr.erase(s->arr.begin() + b, s->arr.begin() + b + cnt);
        if (args.size() > 2) s->arr.insert(s->arr.begin() + b, args.begin() + 2, args.end());
        return Value(Interp::make_array(removed));
    });
    if (key == "includes" || key == "lastIndexOf") return Interp::make_native([self, this, key](Interp&, const Value&, std::vector<Value>& args) {
        auto s = self();
        const Value want = args.empty() ? Value() : args[0];
        for (size_t k = s->arr.size(); k-- > 0;) {
            const bool same = strict_equal(s->arr[k], want) || (key == "includes" && want.type == Value::Number && std::isnan(want.num) && s->arr[k].type == Value::Number && std::isnan(s->arr[k].num));
            if (same) return key == "includes" ? Value(true) : Value((double)k);
        }
        return key == "includes" ? Value(false) : Value(-1.0);
    });
    if (key == "at") return Interp::make_native([self](Interp&, const Value&, std::vector<Value>& args) {
        auto s = self();
        long i = args.empty() ? 0 : (long)Interp::to_number(args[0]);
        if (i < 0) i += (long)s->arr.size();
        return i >= 0 && i < (long)s->arr.size() ? s->arr[(size_t)i] : Value();
    });
    if (key == "find" || key == "findIndex" || key == "findLast" || key == "findLastIndex") return Interp::make_native([self, key](Interp& ip, const Value&, std::vector<Value>& args) {
        auto s = self();
        if (args.empty()) throw JsError("TypeError: callback required");
        const bool last = key == "findLast" || key == "findLastIndex", want_index = key == "findIndex" || key == "findLastIndex";
        const long n = (long)s->arr.size();
        for (long k = last ? n - 1 : 0; last ? k >= 0 : k < n; k += last ? -1 : 1)
            if (Interp::to_boolean(ip.call(args[0], Value(), {s->arr[(size_t)k], Value((double)k), Value(s)}))) return want_index ? Value((double)k) : s->arr[(size_t)k];
        return want_index ? Value(-1.0) : Value();
    });
    if (key == "flat") return Interp::make_native([self](Interp&, const Value&, std::vector<Value>& args) {
        auto s = self();
        const double depth = args.empty() || args[0].is_undefined() ? 1.0 : Interp::to_number(args[0]);
        std::vector<Value> out;
        std::function<void(const std::vector<Value>&, double)> go = [&](const std::vector<Value>& xs, double d) {
            for (auto& v : xs) {
                if (v.type == Value::Obj && v.obj->kind == Object::Array && d >= 1) go(v.obj->arr, d - 1);
                else out.push_back(v);
            }
        };
        go(s->arr, depth);
        return Value(Interp::make_array(out));
    });
    if (key == "flatMap") return Interp::make_native([self](Interp& ip, const Value&, std::vector<Value>& args) {
        auto s = self();
        if (args.empty()) throw JsError("TypeError: callback required");
        std::vector<Value> out;
        for (size_t i = 0; i < s->arr.size(); i++) {
            Value r = ip.call(args[0], Value(), {s->arr[i], Value((double)i), Value(s)});
            if (r.type == Value::Obj && r.obj->kind == Object::Array) out.insert(out.end(), r.obj->arr.begin(), r.obj->arr.end());
            else out.push_back(r);
        }
        return Value(Interp::make_array(out));
    });
    if (key == "reduceRight") return Interp::make_native([self](Interp& ip, const Value&, std::vector<Value>& args) {
        auto s = self();
        if (args.empty()) throw JsError("TypeError: callback required");
        long i = (long)s->arr.size() - 1;
        Value acc;
        if (args.size() > 1) acc = args[1];
        else { if (i < 0) throw JsError("TypeError: reduce of empty array"); acc = s->arr[(size_t)i--]; }
        for (; i >= 0; i--) acc = ip.call(args[0], Value(), {acc, s->arr[(size_t)i], Value((double)i), Value(s)});
        return acc;
    });
    // sort: stable; without a comparator by the elements' strings (10 < 9), undefined last - as JavaScript does
    if (key == "sort") return Interp::make_native([self](Interp& ip, const Value&, std::vector<Value>& args) {
        auto s = self();
        const bool has_cmp = !args.empty() && !args[0].is_undefined();
        Value cmp = has_cmp ? args[0] : Value();
        std::vector<Value> defined, undefined;
        for (auto& v : s->arr) (v.is_undefined() ? undefined : defined).push_back(v);
        std::vector<std::string> keys;
        if (!has_cmp) for (auto& v : defined) keys.push_back(ip.to_string(v));
        std::vector<size_t> order(defined.size());
        for (size_t i = 0; i < order.size(); i++) order[i] = i;
        std::stable_sort(order.begin(), order.end(), [&](size_t a, size_t b) {
            if (!has_cmp) return keys[a] < keys[b];
            const double r = Interp::to_number(ip.call(cmp, Value(), {defined[a], defined[b]}));
            return r < 0;
        });
        std::vector<Value> out;
        for (size_t i : order) out.push_back(defined[i]);
        out.insert(out.end(), undefined.begin(), undefined.end());
        s->arr = out;
        return Value(s);
    });
    if (key == "keys" || key == "values" || key == "entries") return Interp::make_native([self, key](Interp&, const Value&, std::vector<Value>&) {
        // no iterator protocol in this interpreter: the sequence itself, as an array (for-of and spread take it)
        auto s = self();
        std::vector<Value> out;
        for (size_t i = 0; i < s->arr.size(); i++) {
            if (key == "keys") out.push_back(Value((double)i));
            else if (key == "values") out.push_back(s->arr[i]);
            else out.push_back(Value(Interp::make_array({Value((double)i), s->arr[i]})));
        }
        return Value(Interp::make_array(out));
    });
    if (key == "map" || key == "forEach" || key == "filter" || key == "reduce" || key == "some" || key == "every")
        return Interp::make_native([self, key](Interp& ip, const Value&, std::vector<Value>& args) {
            auto s = self();
            if (args.empty()) throw JsError("TypeError: callback required");
            std::vector<Value> out;
            Value acc;
            size_t start = 0;
            if (key == "reduce") {
                if (args.size() > 1) acc = args[1];
                else { if (s->arr.empty()) throw JsError("TypeError: reduce of empty array"); acc = s->arr[0]; start = 1; }
            }
            for (size_t i = start; i < s->arr.size(); i++) {
                if (key == "reduce") { acc = ip.call(args[0], Value(), {acc, s->arr[i], Value((double)i), Value(s)}); continue; }
                Value r = ip.call(args[0], Value(), {s->arr[i], Value((double)i), Value(s)});
                if (key == "map") out.push_back(r);
                else if (key == "filter") { if (Interp::to_boolean(r)) out.push_back(s->arr[i]); }
                else if (key == "some") { if (Interp::to_boolean(r)) return Value(true); }
                else if (key == "every") { if (!Interp::to_boolean(r)) return Value(false); }
            }
            if (key == "map" || key == "filter") return Value(Interp::make_array(out));
            if (key == "reduce") return acc;
            if (key == "some") return Value(false);
            if (key == "every") return Value(true);
            return Value();
        });
    (void)n;
    return Value();
}

// Methods of string values (ASCII semantics: metric scripts build names and labels, not text)
Value Evaluator::string_method(const std::string& str, const std::string& key) {
    auto arg_str = [](Interp& ip, std::vector<Value>& a, size_t i) { return i < a.size() ? ip.to_string(a[i]) : std::string("undefined"); };
    auto clamp_index = [](double d, long len) { long i = std::isnan(d) ? 0 : (long)std::max(-1e15, std::min(1e15, d)); if (i < 0) i = std::max(0L, len + i); return std::min(i, len); };
    const std::string s = str;
    if (key == "charAt") return Value(Interp::make_native([s](Interp&, const Value&, std::vector<Value>& a) { long i = a.empty() ? 0 : (long)Interp::to_number(a[0]); return Value(i >= 0 && i < (long)s.size() ? std::string(1, s[(size_t)i]) : std::string()); }));
    if (key == "charCodeAt") return Value(Interp::make_native([s](Interp&, const Value&, std::vector<Value>& a) { long i = a.empty() ? 0 : (long)Interp::to_number(a[0]); return Value(i >= 0 && i < (long)s.size() ? (double)(unsigned char)s[(size_t)i] : NAN); }));
    if (key == "at") return Value(Interp::make_native([s](Interp&, const Value&, std::vector<Value>& a) { long i = a.empty() ? 0 : (long)Interp::to_number(a[0]); if (i < 0) i += (long)s.size(); return i >= 0 && i < (long)s.size() ? Value(std::string(1, s[(size_t)i])) : Value(); }));
    if (key == "indexOf" || key == "lastIndexOf" || key == "includes" || key == "startsWith" || key == "endsWith")
        return Value(Interp::make_native([s, key, arg_str](Interp& ip, const Value&, std::vector<Value>& a) {
            const std::string what = arg_str(ip, a, 0);
            if (key == "startsWith") return Value(s.compare(0, what.size(), what) == 0);
            if (key == "endsWith") return Value(s.size() >= what.size() && s.compare(s.size() - what.size(), what.size(), what) == 0);
            const size_t at = key == "lastIndexOf" ? s.rfind(what) : s.find(what, a.size() > 1 ? (size_t)std::max(0.0, Interp::to_number(a[1])) : 0);
            if (key == "includes") return Value(at != std::string::npos);
            return Value(at == std::string::npos ? -1.0 : (double)at);
        }));
    if (key == "slice" || key == "substring" || key == "substr")
        return Value(Interp::make_native([s, key, clamp_index](Interp&, const Value&, std::vector<Value>& a) {
            const long len = (long)s.size();
            const double d0 = a.empty() ? 0.0 : Interp::to_number(a[0]), d1 = a.size() > 1 && !a[1].is_undefined() ? Interp::to_number(a[1]) : (double)len;
            long b, e;
            if (key == "slice") { b = clamp_index(d0, len); e = clamp_index(d1, len); }
            else if (key == "substr") { b = clamp_index(d0, len); e = a.size() > 1 && !a[1].is_undefined() ? std::min(len, b + std::max(0L, (long)d1)) : len; }
            else { b = std::min(len, std::max(0L, std::isnan(d0) ? 0L : (long)d0)); e = std::min(len, std::max(0L, std::isnan(d1) ? 0L : (long)d1)); if (b > e) std::swap(b, e); }
            return Value(e > b ? s.substr((size_t)b, (size_t)(e - b)) : std::string());
        }));
    if (key == "toUpperCase" || key == "toLowerCase")
        return Value(Interp::make_native([s, key](Interp&, const Value&, std::vector<Value>&) { std::string o = s; for (auto& c : o) c = (char)(key == "toUpperCase" ? toupper((unsigned char)c) : tolower((unsigned char)c)); return Value(o); }));
    if (key == "trim" || key == "trimStart" || key == "trimEnd")
        return Value(Interp::make_native([s, key](Interp&, const Value&, std::vector<Value>&) {
            size_t b = 0, e = s.size();
            if (key != "trimEnd") while (b < e && isspace((unsigned char)s[b])) b++;
            if (key != "trimStart") while (e > b && isspace((unsigned char)s[e - 1])) e--;
            return Value(s.substr(b, e - b));
        }));
    if (key == "repeat") return Value(Interp::make_native([s](Interp&, const Value&, std::vector<Value>& a) {
        const double n = a.empty() ? 0 : Interp::to_number(a[0]);
        if (n < 0 || n * (double)s.size() > 1e7) throw JsError("RangeError: invalid count value");
        std::string o;
        for (long i = 0; i < (long)n; i++) o += s;
        return Value(o);
    }));
    if (key == "padStart" || key == "padEnd") return Value(Interp::make_native([s, key](Interp& ip, const Value&, std::vector<Value>& a) {
        const long want = a.empty() ? 0 : (long)Interp::to_number(a[0]);
        const std::string fill = a.size() > 1 && !a[1].is_undefined() ? ip.to_string(a[1]) : " ";
        if (want <= (long)s.size() || fill.empty() || want > 10000000) return Value(s);
        std::string pad;
        while ((long)(pad.size() + s.size()) < want) pad += fill;
        pad.resize((size_t)want - s.size());
        return Value(key == "padStart" ? pad + s : s + pad);
    }));
    if (key == "concat") return Value(Interp::make_native([s](Interp& ip, const Value&, std::vector<Value>& a) { std::string o = s; for (auto& v : a) o += ip.to_string(v); return Value(o); }));
    if (key == "split") return Value(Interp::make_native([s](Interp& ip, const Value&, std::vector<Value>& a) {
        std::vector<Value> out;
        if (a.empty() || a[0].is_undefined()) return Value(Interp::make_array({Value(s)}));
        const std::string sep = ip.to_string(a[0]);
        if (sep.empty()) { for (char c : s) out.push_back(Value(std::string(1, c))); return Value(Interp::make_array(out)); }
        size_t from = 0;
        for (;;) {
            const size_t at = s.find(sep, from);
            out.push_back(Value(s.substr(from, at == std::string::npos ? std::string::npos : at - from)));
            if (at == std::string::npos) break;
            from = at + sep.size();
        }
        return Value(Interp::make_array(out));
    }));
    if (key == "replace" || key == "replaceAll") return Value(Interp::make_native([s, key](Interp& ip, const Value&, std::vector<Value>& a) {
        // string patterns only (there are no regular expressions in this interpreter)
        const std::string what = a.empty() ? "undefined" : ip.to_string(a[0]), with = a.size() > 1 ? ip.to_string(a[1]) : "undefined";
        std::string o = s;
        size_t from = 0;
        while (!what.empty()) {
            const size_t at = o.find(what, from);
            if (at == std::string::npos) break;
            o.replace(at, what.size(), with);
            from = at + with.size();
            if (key == "replace") break;
        }
        return Value(o);
    }));
    if (key == "toString" || key == "valueOf") return Value(Interp::make_native([s](Interp&, const Value&, std::vector<Value>&) { return Value(s); }));
    return Value();
}

// Methods of number values
Value Evaluator::number_method(double num, const std::string& key) {
    if (key == "toFixed") return Value(Interp::make_native([num](Interp& ip, const Value&, std::vector<Value>& a) {
        const int digits = a.empty() ? 0 : (int)Interp::to_number(a[0]);
        if (digits < 0 || digits > 100) throw JsError("RangeError: toFixed() digits argument must be between 0 and 100");
        if (!std::isfinite(num) || std::fabs(num) >= 1e21) return Value(ip.to_string(Value(num)));
        char buf[160];
        snprintf(buf, sizeof buf, "%.*f", digits, num);
        return Value(std::string(buf));
    }));
    if (key == "toPrecision") return Value(Interp::make_native([num](Interp& ip, const Value&, std::vector<Value>& a) {
        if (a.empty() || a[0].is_undefined() || !std::isfinite(num)) return Value(ip.to_string(Value(num)));
        const int p = (int)Interp::to_number(a[0]);
        if (p < 1 || p > 100) throw JsError("RangeError: toPrecision() argument must be between 1 and 100");
        char buf[160];
        snprintf(buf, sizeof buf, "%.*e", p - 1, num);
        const int e = atoi(strchr(buf, 'e') + 1);
        if (e < -6 || e >= p) {                                        // exponential form, "e+N" without zero padding
            std::string m(buf, strchr(buf, 'e'));
            return Value(m + "e" + (e >= 0 ? "+" : "-") + std::to_string(std::abs(e)));
        }
        snprintf(buf, sizeof buf, "%.*f", std::max(0, p - 1 - e), num);
        return Value(std::string(buf));
    }));
    if (key == "toString") return Value(Interp::make_native([num](Interp& ip, const Value&, std::vector<Value>& a) {
        const int radix = a.empty() || a[0].is_undefined() ? 10 : (int)Interp::to_number(a[0]);
        if (radix == 10 || !std::isfinite(num) || num != std::floor(num)) return Value(ip.to_string(Value(num)));
        if (radix < 2 || radix > 36) throw JsError("RangeError: toString() radix must be between 2 and 36");
        std::string o;
        double v = std::fabs(num);
        if (v == 0) o = "0";
        while (v >= 1) { const int d = (int)std::fmod(v, radix); o.insert(o.begin(), (char)(d < 10 ? '0' + d : 'a' + d - 10)); v = std::floor(v / radix); }
        return Value(num < 0 ? "-" + o : o);
    }));
    if (key == "valueOf") return Value(Interp::make_native([num](Interp&, const Value&, std::vector<Value>&) { return Value(num); }));
    return Value();
}

Value Evaluator::get_prop(const Node* n, const Value& objv, const std::string& key) {
    if (objv.type == Value::Undefined || objv.type == Value::Null)
        fail(n, "TypeError: cannot read property '" + key + "' of " + (objv.type == Value::Null ? "null" : "undefined"));
    if (objv.type == Value::String) {
        if (key == "length") return Value((double)objv.str.size());
        size_t idx;
        if (array_index(key, idx) && idx < objv.str.size()) return Value(std::string(1, objv.str[idx]));
        return string_method(objv.str, key);
    }
    if (objv.type == Value::Number) return number_method(objv.num, key);
    if (objv.type != Value::Obj) return Value();
    const ObjP& o = objv.obj;
    if (o->get_hook) {
        Value out;
        if (o->get_hook(in, key, out)) return out;
    }
    if (o->kind == Object::Array) {
        if (key == "length") return Value((double)o->arr.size());
        size_t idx;
        if (array_index(key, idx)) return idx < o->arr.size() ? o->arr[idx] : Value();
        auto it = o->props.find(key);
        if (it != o->props.end()) return it->second;
        return array_method(n, o, key);
    }
    if (o->kind == Object::Function && key == "prototype") return prototype_of_function(o);
    // own property, own accessor, then the same along the prototype chain (an instance's methods live on C.prototype, a class's
    // inherited statics on its parent class)
    for (const Object* c = o.get(); c; c = c->proto.get()) {
        auto it = c->props.find(key);
        if (it != c->props.end()) return it->second;
        auto g = c->props.find("$get$" + key);
        if (g != c->props.end()) { std::vector<Value> none; return call_function(n, g->second, objv, none); }
    }
    if (o->kind == Object::Function && o->def && key == "name") return Value(o->def->name);
    if (o->kind == Object::Function && o->def && key == "length") {
        size_t n_params = 0;
        for (size_t i = 0; i < o->def->params.size() && !o->def->defaults[i]; i++) n_params++;       // parameters before the first default
        return Value((double)n_params);
    }
    if ((o->kind == Object::Function || o->kind == Object::Native) && key == "call") {
        Value fn = objv;
        return Value(Interp::make_native([fn](Interp& ip, const Value&, std::vector<Value>& args) {
            Value self = args.empty() ? Value() : args[0];
            std::vector<Value> rest(args.size() > 1 ? args.begin() + 1 : args.end(), args.end());
            return ip.call(fn, self, rest);
        }));
    }
    if ((o->kind == Object::Function || o->kind == Object::Native) && key == "apply") {
        Value fn = objv;
        return Value(Interp::make_native([fn](Interp& ip, const Value&, std::vector<Value>& args) {
            Value self = args.empty() ? Value() : args[0];
            std::vector<Value> rest;
            if (args.size() > 1 && args[1].type == Value::Obj && args[1].obj->kind == Object::Array) rest = args[1].obj->arr;
            return ip.call(fn, self, rest);
        }));
    }
    if ((o->kind == Object::Function || o->kind == Object::Native) && key == "bind") {
        Value fn = objv;
        return Value(Interp::make_native([fn](Interp&, const Value&, std::vector<Value>& args) {
            Value self = args.empty() ? Value() : args[0];
            std::vector<Value> bound(args.size() > 1 ? args.begin() + 1 : args.end(), args.end());
            return Value(Interp::make_native([fn, self, bound](Interp& ip, const Value&, std::vector<Value>& more) {
                std::vector<Value> all = bound;
                all.insert(all.end(), more.begin(), more.end());
                return ip.call(fn, self, all);
            }));
        }));
    }
    if (key == "hasOwnProperty") {                            // o.hasOwnProperty(k); also reached through Object.prototype.hasOwnProperty.call(o, k)
        return Value(Interp::make_native([](Interp& ip, const Value& self, std::vector<Value>& args) {
            if (self.type != Value::Obj) return Value(false);
            const std::string k = args.empty() ? "undefined" : ip.to_string(args[0]);
            for (auto& own : Interp::own_keys(self)) if (own == k) return Value(true);
            return Value(false);
        }));
    }
    if (o.get() == in.builtin_object.get() && key == "prototype") return Value(Interp::make_object());    // Object.prototype.hasOwnProperty.call(...)
    return Value();
}

void Evaluator::set_prop(const Node* n, const Value& objv, const std::string& key, const Value& v) {
    if (objv.type != Value::Obj) fail(n, "TypeError: cannot set property '" + key + "' on a non-object");
    const ObjP& o = objv.obj;
    if (o->set_hook && o->set_hook(in, key, v)) return;
    if (o->kind == Object::Array) {
        if (key == "length") {
            double d = Interp::to_number(v);
            if (d < 0 || d > 1e7) fail(n, "RangeError: invalid array length");
            o->arr.resize((size_t)d);
            return;
        }
        size_t idx;
        if (array_index(key, idx)) {
            if (idx > 10000000) fail(n, "RangeError: array index too large");
            if (idx >= o->arr.size()) o->arr.resize(idx + 1);
            o->arr[idx] = v;
            return;
        }
    }
    if (!o->props.count(key)) {
        // a setter of the object itself (object literal, static), of its class or of a parent; a getter without a setter swallows the
        // assignment (sloppy mode)
        for (const Object* c = o.get(); c; c = c->proto.get()) {
            auto st = c->props.find("$set$" + key);
            if (st != c->props.end()) { std::vector<Value> arg{v}; call_function(n, st->second, objv, arg); return; }
            if (c->props.count("$get$" + key)) return;
            if (c != o.get() && c->props.count(key)) break;         // a plain inherited value: shadowed by an own property
        }
    }
    o->props[key] = v;
}

bool Evaluator::strict_equal(const Value& a, const Value& b) {
    if (a.type != b.type) return false;
    switch (a.type) {
        case Value::Undefined: case Value::Null: return true;
        case Value::Bool: return a.b == b.b;
        case Value::Number: return a.num == b.num;
        case Value::String: return a.str == b.str;
        case Value::Obj: return a.obj == b.obj;
    }
    return false;
}

bool Evaluator::loose_equal(const Value& a, const Value& b) {
    if (a.type == b.type) return strict_equal(a, b);
    bool an = a.type == Value::Undefined || a.type == Value::Null;
    bool bn = b.type == Value::Undefined || b.type == Value::Null;
    if (an || bn) return an && bn;
    if (a.type == Value::Obj || b.type == Value::Obj) return false;
    return Interp::to_number(a) == Interp::to_number(b);
}

Value Evaluator::binary_op(const Node* n, const std::string& op, const Value& a, const Value& b) {
    if (is_host(a) || is_host(b)) {
        if (op == "===" ) return Value(strict_equal(a, b));
        if (op == "!==") return Value(!strict_equal(a, b));
        if (!in.host_operator) fail(n, "TypeError: no operator overloads installed");
        return in.host_operator(in, op, a, b);
    }
    if (op != "===" && op != "!==" && op != "==" && op != "!=" && op != "in" && op != "instanceof") {
        // an object with its own valueOf / toString takes part as that primitive
        Value pa = to_primitive(n, a), pb = to_primitive(n, b);
        if (pa.type != Value::Obj || pb.type != Value::Obj) {
            if ((a.type == Value::Obj && pa.type != Value::Obj) || (b.type == Value::Obj && pb.type != Value::Obj)) return binary_op(n, op, pa, pb);
        }
    }
    if (op == "+") {
        bool as = a.type == Value::String || (a.type == Value::Obj), bs = b.type == Value::String || (b.type == Value::Obj);
        if (as || bs) return Value(in.to_string(a) + in.to_string(b));
        return Value(Interp::to_number(a) + Interp::to_number(b));
    }
    if (op == "===") return Value(strict_equal(a, b));
    if (op == "!==") return Value(!strict_equal(a, b));
    if (op == "==") return Value(loose_equal(a, b));
    if (op == "!=") return Value(!loose_equal(a, b));
    if (a.type == Value::String && b.type == Value::String) {
        if (op == "<") return Value(a.str < b.str);
        if (op == ">") return Value(a.str > b.str);
        if (op == "<=") return Value(a.str <= b.str);
        if (op == ">=") return Value(a.str >= b.str);
    }
    if (op == "in") {
        if (b.type != Value::Obj) fail(n, "TypeError: 'in' needs an object");
        std::string k = key_of(a);
        size_t idx;
        if (b.obj->kind == Object::Array && array_index(k, idx)) return Value(idx < b.obj->arr.size());
        return Value(holder_of(b.obj.get(), k) != nullptr || holder_of(b.obj.get(), "$get$" + k) != nullptr);
    }
    if (op == "instanceof") {
        if (b.type != Value::Obj || (b.obj->kind != Object::Function && b.obj->kind != Object::Native)) fail(n, "TypeError: right-hand side of 'instanceof' is not callable");
        if (a.type != Value::Obj) return Value(false);
        for (const Object* c = a.obj->klass.get(); c; c = c->klass.get())                  // error kinds
            if (c == b.obj.get()) return Value(true);
        if (b.obj->kind == Object::Function) {
            Value bp = prototype_of_function(b.obj);
            for (const Object* c = a.obj->proto.get(); c && bp.type == Value::Obj; c = c->proto.get())
                if (c == bp.obj.get()) return Value(true);
        }
        for (const Object* c = a.obj->proto.get(); c; c = c->proto.get())                   // class E extends Error: the chain reaches a native kind through the class
            for (const Object* k = c->klass.get(); k; k = k->klass.get())
                if (k == b.obj.get()) return Value(true);
        if (b.obj.get() == in.builtin_array.get()) return Value(a.obj->kind == Object::Array);
        if (b.obj.get() == in.builtin_object.get()) return Value(true);
        if (b.obj.get() == in.builtin_function.get()) return Value(a.obj->kind == Object::Function || a.obj->kind == Object::Native);
        return Value(false);
    }
    double x = Interp::to_number(a), y = Interp::to_number(b);
    if (op == "-") return Value(x - y);
    if (op == "*") return Value(x * y);
    if (op == "/") return Value(x / y);
    if (op == "%") return Value(std::fmod(x, y));
    if (op == "**") return Value(std::pow(x, y));
    if (op == "<") return Value(x < y);
    if (op == ">") return Value(x > y);
    if (op == "<=") return Value(x <= y);
    if (op == ">=") return Value(x >= y);
    auto i32 = [](double d) -> int32_t { return std::isfinite(d) ? (int32_t)(int64_t)d : 0; };
    if (op == "&") return Value((double)(i32(x) & i32(y)));
    if (op == "|") return Value((double)(i32(x) | i32(y)));
    if (op == "^") return Value((double)(i32(x) ^ i32(y)));
    if (op == "<<") return Value((double)(i32(x) << (i32(y) & 31)));
    if (op == ">>") return Value((double)(i32(x) >> (i32(y) & 31)));
    if (op == ">>>") return Value((double)((uint32_t)i32(x) >> (i32(y) & 31)));
    fail(n, "unsupported operator '" + op + "'");
}

void Evaluator::assign_to(const Node* target, EnvP& env, const Value& v) {
    if (target->k == K::Ident) set_var(target, env, target->s, v);
    else if (target->k == K::Member) set_prop(target, eval(target->kids[0].get(), env), target->s, v);
    else if (target->k == K::Index) {
        Value o = eval(target->kids[0].get(), env);
        Value k = eval(target->kids[1].get(), env);
        set_prop(target, o, key_of(k), v);
    } else fail(target, "invalid assignment target");
}

Value Evaluator::call_function(const Node* n, const Value& fn, const Value& self, std::vector<Value>& args) {
    if (fn.type != Value::Obj || (fn.obj->kind != Object::Function && fn.obj->kind != Object::Native))
        fail(n, "TypeError: value is not a function");
    if (fn.obj->kind == Object::Native) return fn.obj->native(in, self, args);
    if (++in.call_depth > 200) { in.call_depth--; fail(n, "RangeError: call stack too deep"); }
    const FuncDef* def = fn.obj->def;
    EnvP env = new_env(fn.obj->env, true);
    env->this_val = fn.obj->is_arrow ? fn.obj->bound_this : self;
    if (!fn.obj->is_arrow && !def->strict && (self.type == Value::Undefined || self.type == Value::Null)) {
        auto g = in.globals->vars.find("globalThis");            // sloppy mode (scripts are not modules): a plain call sees the global object
        if (g != in.globals->vars.end()) env->this_val = g->second;
    }
    for (size_t i = 0; i < def->params.size(); i++) {
        Value v = i < args.size() ? args[i] : Value();
        if (v.is_undefined() && i < def->defaults.size() && def->defaults[i]) v = eval(def->defaults[i].get(), env);   // earlier parameters are in scope
        env->vars[def->params[i]] = v;
    }
    if (!def->rest.empty())
        env->vars[def->rest] = Value(Interp::make_array(args.size() > def->params.size() ? std::vector<Value>(args.begin() + def->params.size(), args.end()) : std::vector<Value>()));
    if (!def->is_arrow) env->vars["arguments"] = Value(Interp::make_array(args));
    hoist(*def, env);
    for (auto& d : def->prologue) exec(d.get(), env);
    Value result;
    try {
        if (def->expr_body) result = eval(def->body.get(), env);
        else {
            Value saved_ret = ret;
            ret = Value();
            Comp c = exec_block(def->body->kids, env);
            if (c == Comp::Return) result = ret;
            ret = saved_ret;
        }
    } catch (...) {
        in.call_depth--;
        throw;
    }
    in.call_depth--;
    return result;
}

Value Evaluator::eval(const Node* n, EnvP& env) {
    tick(n);
    switch (n->k) {
        case K::Num: return Value(n->num);
        case K::Str: return Value(n->s);
        case K::Literal:
            if (n->s == "bool") return Value(n->flag);
            if (n->s == "null") return Value::null();
            return Value();
        case K::This: return env->this_val;
        case K::Ident: return get_var(n, env, n->s);
        case K::ArrayLit: {
            std::vector<Value> items;
            for (auto& k : n->kids) push_arg(k.get(), env, items);
            return Value(Interp::make_array(items));
        }
        case K::ObjectLit: {
            auto o = Interp::make_object();
            for (size_t i = 0; i + 1 < n->kids.size(); i += 2) {
                if (n->kids[i]->k != K::Str) {                                      // {[expr]: value}: the key is evaluated first
                    const std::string key = key_of(eval(n->kids[i].get(), env));
                    o->props[key] = eval(n->kids[i + 1].get(), env);
                    continue;
                }
                if (n->kids[i + 1]->k == K::Assign && n->kids[i + 1]->flag) fail(n->kids[i + 1].get(), "SyntaxError: invalid shorthand property initializer");
                Value v = eval(n->kids[i + 1].get(), env);
                if (n->kids[i]->s == "__proto__" && n->kids[i + 1]->k != K::Func) {   // {__proto__: base}: the prototype, not a property
                    if (v.type == Value::Obj) o->proto = v.obj;
                    continue;
                }
                if (n->kids[i]->s == "..." && n->kids[i + 1]->k != K::Str) {      // {...base}
                    if (v.type == Value::Obj) {
                        if (v.obj->kind == Object::Array) for (size_t k = 0; k < v.obj->arr.size(); k++) o->props[std::to_string(k)] = v.obj->arr[k];
                        for (auto& kv : v.obj->props) o->props[kv.first] = kv.second;
                    }
                    continue;
                }
                o->props[n->kids[i]->s] = v;
            }
            return Value(o);
        }
        case K::Spread: fail(n, "SyntaxError: '...' is only valid in a call, an array literal or an object literal");
        case K::Class: return make_class(n, env);
        case K::Func: {
            if (!n->fn->is_arrow && !n->fn->name.empty() && n->fn->name != "constructor") {      // const f = function fact(n) { ... fact(n - 1) ... }
                EnvP named = new_env(env, false);
                Value f = make_closure(n->fn.get(), named);
                named->vars[n->fn->name] = f;
                return f;
            }
            return make_closure(n->fn.get(), env);
        }
        case K::OptChain: {
            const bool outer = chain_short;
            chain_short = false;
            Value v = eval(n->kids[0].get(), env);
            if (chain_short) v = Value();
            chain_short = outer;
            return v;
        }
        case K::Member: {
            Value o = eval(n->kids[0].get(), env);
            if (chain_short) return Value();
            if (n->kids[0]->k == K::Ident && n->kids[0]->s == "super" && o.type == Value::Obj) {
                if (const Object* h = holder_of(o.obj.get(), "$get$" + n->s)) {
                    std::vector<Value> none;
                    return call_function(n, h->props.find("$get$" + n->s)->second, env->this_val, none);
                }
            }
            if (n->flag && (o.type == Value::Undefined || o.type == Value::Null)) { chain_short = true; return Value(); }
            return get_prop(n, o, n->s);
        }
        case K::Index: {
            Value o = eval(n->kids[0].get(), env);
            if (chain_short) return Value();
            if (n->flag && (o.type == Value::Undefined || o.type == Value::Null)) { chain_short = true; return Value(); }
            Value k = eval(n->kids[1].get(), env);
            return get_prop(n, o, key_of(k));
        }
        case K::Call: {
            const Node* callee = n->kids[0].get();
            Value self, fn;
            if (callee->k == K::Member) {
                self = eval(callee->kids[0].get(), env);
                if (chain_short) return Value();
                if (callee->flag && (self.type == Value::Undefined || self.type == Value::Null)) { chain_short = true; return Value(); }
                fn = get_prop(callee, self, callee->s);
                if (n->flag && (fn.type == Value::Undefined || fn.type == Value::Null)) { chain_short = true; return Value(); }
                if (fn.is_undefined()) fail(n, "TypeError: '" + callee->s + "' is not a function");
            } else if (callee->k == K::Index) {
                self = eval(callee->kids[0].get(), env);
                if (chain_short) return Value();
                if (callee->flag && (self.type == Value::Undefined || self.type == Value::Null)) { chain_short = true; return Value(); }
                fn = get_prop(callee, self, key_of(eval(callee->kids[1].get(), env)));
                if (n->flag && (fn.type == Value::Undefined || fn.type == Value::Null)) { chain_short = true; return Value(); }
            } else {
                fn = eval(callee, env);
                if (chain_short) return Value();
                if (n->flag && (fn.type == Value::Undefined || fn.type == Value::Null)) { chain_short = true; return Value(); }
            }
            // super(...) and super.method(...) run with the current `this`
            if ((callee->k == K::Ident && callee->s == "super") || (callee->k == K::Member && callee->kids[0]->k == K::Ident && callee->kids[0]->s == "super")) self = env->this_val;
            std::vector<Value> args;
            for (size_t i = 1; i < n->kids.size(); i++) push_arg(n->kids[i].get(), env, args);
            if (fn.type != Value::Obj) fail(n, "TypeError: " + (callee->k == K::Ident ? "'" + callee->s + "'" : std::string("expression")) + " is not a function");
            if (fn.obj->is_class && !(callee->k == K::Ident && callee->s == "super")) fail(n, "TypeError: class constructors must be invoked with 'new'");
            if (callee->k == K::Ident && callee->s == "super" && fn.obj->kind == Object::Native && self.type == Value::Obj) {
                Value made = call_function(n, fn, self, args);          // class MyErr extends Error: the native parent builds {message, name}
                if (made.type == Value::Obj) for (auto& kv : made.obj->props) self.obj->props[kv.first] = kv.second;
                return Value();
            }
            return call_function(n, fn, self, args);
        }
        case K::New: {
            Value ctor = eval(n->kids[0].get(), env);
            std::vector<Value> args;
            for (size_t i = 1; i < n->kids.size(); i++) push_arg(n->kids[i].get(), env, args);
            if (ctor.type != Value::Obj) fail(n, "TypeError: not a constructor");
            if (ctor.obj->kind == Object::Native) return ctor.obj->native(in, Value(), args);
            Value self(Interp::make_object());
            {
                Value pr = prototype_of_function(ctor.obj);          // methods are found through the chain, the instance owns its fields only
                if (pr.type == Value::Obj) self.obj->proto = pr.obj;
            }
            Value r = call_function(n, ctor, self, args);
            return r.type == Value::Obj ? r : self;
        }
        case K::Unary: {
            if (n->s == "delete" && (n->kids[0]->k == K::Member || n->kids[0]->k == K::Index)) {       // delete o.k / delete o[k] / delete a[i]
                const Node* target = n->kids[0].get();
                Value o = eval(target->kids[0].get(), env);
                const std::string key = target->k == K::Member ? target->s : key_of(eval(target->kids[1].get(), env));
                if (o.type == Value::Obj) {
                    size_t idx;
                    if (o.obj->kind == Object::Array && array_index(key, idx)) { if (idx < o.obj->arr.size()) o.obj->arr[idx] = Value(); }   // a hole reads as undefined
                    else o.obj->props.erase(key);
                }
                return Value(true);
            }
            if (n->s == "typeof") {
                Value v;
                if (n->kids[0]->k == K::Ident && !lookup_env(env.get(), n->kids[0]->s) && !in.global_getters.count(n->kids[0]->s)) return Value("undefined");
                v = eval(n->kids[0].get(), env);
                switch (v.type) {
                    case Value::Undefined: return Value("undefined");
                    case Value::Null: return Value("object");
                    case Value::Bool: return Value("boolean");
                    case Value::Number: return Value("number");
                    case Value::String: return Value("string");
                    case Value::Obj: return Value((v.obj->kind == Object::Function || v.obj->kind == Object::Native) ? "function" : "object");
                }
            }
            Value v = eval(n->kids[0].get(), env);
            if (n->s == "void") return Value();
            if (n->s == "delete") return Value(true);      // (a delete of a property is handled before the operand is evaluated, below)
            if (n->s == "!") return Value(!Interp::to_boolean(v));
            if (is_host(v)) {
                if (!in.host_operator) fail(n, "TypeError: no operator overloads installed");
                if (n->s == "-") return in.host_operator(in, "neg", v, Value());
                if (n->s == "+") fail(n, "TypeError: no overloaded operator pos");       // number.js overloads neg, not pos: the reference's engine refuses +x on a number object
                fail(n, "TypeError: unsupported unary operator on number object");
            }
            if (n->s == "-") return Value(-Interp::to_number(v));
            if (n->s == "+") return Value(Interp::to_number(v));
            if (n->s == "~") return Value((double)(~(int32_t)(int64_t)Interp::to_number(v)));
            fail(n, "unsupported unary operator");
        }
        case K::Update: {
            const Node* target = n->kids[0].get();
            Value old = eval(target, env);
            Value one(1.0);
            Value nv = is_host(old) ? binary_op(n, n->s == "++" ? "+" : "-", old, one)
                                    : Value(Interp::to_number(old) + (n->s == "++" ? 1.0 : -1.0));
            assign_to(target, env, nv);
            return n->flag ? nv : (is_host(old) ? old : Value(Interp::to_number(old)));
        }
        case K::Binary: {
            Value a = eval(n->kids[0].get(), env);
            Value b = eval(n->kids[1].get(), env);
            return binary_op(n, n->s, a, b);
        }
        case K::Logical: {
            Value a = eval(n->kids[0].get(), env);
            if (n->s == "&&") return Interp::to_boolean(a) ? eval(n->kids[1].get(), env) : a;
            if (n->s == "||") return Interp::to_boolean(a) ? a : eval(n->kids[1].get(), env);
            return (a.type == Value::Undefined || a.type == Value::Null) ? eval(n->kids[1].get(), env) : a;
        }
        case K::ObjectRest: {
            Value src = eval(n->kids[0].get(), env);
            auto o = Interp::make_object();
            if (src.type == Value::Obj) {
                const std::string skip = '\x1f' + n->s;         // "\x1fa\x1fb\x1f"
                for (auto& kv : src.obj->props)
                    if (skip.find('\x1f' + kv.first + '\x1f') == std::string::npos && kv.first.compare(0, 5, "$get$") != 0 && kv.first.compare(0, 5, "$set$") != 0) o->props[kv.first] = kv.second;
            }
            return Value(o);
        }
        case K::Assign: {
            const Node* target = n->kids[0].get();
            if (n->s == "=" && (target->k == K::ArrayLit || target->k == K::ObjectLit)) {      // [a, b] = value;  ({a, b: [c]} = value)
                Value v = eval(n->kids[1].get(), env);
                destructure_assign(target, env, v);
                return v;
            }
            if (n->s == "=") {
                // evaluate target object before the value, as JavaScript does
                if (target->k == K::Ident) {
                    Value v = eval(n->kids[1].get(), env);
                    set_var(target, env, target->s, v);
                    return v;
                }
                Value o = eval(target->kids[0].get(), env);
                std::string key = target->k == K::Member ? target->s : key_of(eval(target->kids[1].get(), env));
                Value v = eval(n->kids[1].get(), env);
                set_prop(target, o, key, v);
                return v;
            }
            Value old = eval(target, env);
            Value rhs = eval(n->kids[1].get(), env);
            Value nv = binary_op(n, n->s.substr(0, n->s.size() - 1), old, rhs);
            assign_to(target, env, nv);
            return nv;
        }
        case K::Cond:
            return Interp::to_boolean(eval(n->kids[0].get(), env)) ? eval(n->kids[1].get(), env) : eval(n->kids[2].get(), env);
        case K::Comma: {
            Value v;
            for (auto& k : n->kids) v = eval(k.get(), env);
            return v;
        }
        default: fail(n, "internal: statement node in expression position");
    }
}

Comp Evaluator::exec_block(const std::vector<NodeP>& body, EnvP& env) {
    for (auto& s : body) {
        Comp c = exec(s.get(), env);
        if (c != Comp::Normal) return c;
    }
    return Comp::Normal;
}

Comp Evaluator::exec(const Node* n, EnvP& env) {
    tick(n);
    switch (n->k) {
        case K::Empty: return Comp::Normal;
        case K::ExprStmt: completion = eval(n->kids[0].get(), env); return Comp::Normal;
        case K::VarDecl: {
            for (auto& d : n->kids) {
                if (n->s == "var") {
                    if (!d->kids.empty()) set_var(d.get(), env, d->s, eval(d->kids[0].get(), env));
                } else {
                    Value v = d->kids.empty() ? Value() : eval(d->kids[0].get(), env);
                    env->vars[d->s] = v;
                    env->is_const[d->s] = (n->s == "const");
                }
            }
            return Comp::Normal;
        }
        case K::FuncDecl: return Comp::Normal;  // hoisted
        case K::Class: {
            env->vars[n->s] = make_class(n, env);
            return Comp::Normal;
        }
        case K::Switch: {
            Value d = eval(n->kids[0].get(), env);
            EnvP inner = new_env(env, false);
            size_t start = n->kids.size(), fallback = n->kids.size();
            for (size_t i = 1; i < n->kids.size() && start == n->kids.size(); i++) {
                const Node* c = n->kids[i].get();
                if (c->kids[0]->k == K::Empty) { fallback = i; continue; }
                if (strict_equal(d, eval(c->kids[0].get(), inner))) start = i;
            }
            if (start == n->kids.size()) start = fallback;
            for (size_t i = start; i < n->kids.size(); i++) {        // fall through until a break
                const Node* c = n->kids[i].get();
                for (size_t k = 1; k < c->kids.size(); k++) {
                    Comp r = exec(c->kids[k].get(), inner);
                    if (r == Comp::Break && jump_label.empty()) return Comp::Normal;
                    if (r != Comp::Normal) return r;
                }
            }
            return Comp::Normal;
        }
        case K::Return:
            ret = n->kids.empty() ? Value() : eval(n->kids[0].get(), env);
            return Comp::Return;
        case K::If:
            if (Interp::to_boolean(eval(n->kids[0].get(), env))) return exec(n->kids[1].get(), env);
            if (n->kids.size() > 2) return exec(n->kids[2].get(), env);
            return Comp::Normal;
        case K::Block: {
            EnvP inner = new_env(env, false);
            // block-level function declarations are visible in the block
            for (auto& s : n->kids)
                if (s->k == K::FuncDecl && !lookup_env(env.get(), s->s)) inner->vars[s->s] = make_closure(s->fn.get(), inner);
            return exec_block(n->kids, inner);
        }
        case K::For: {
            const std::vector<std::string> mine = take_loop_labels();
            EnvP loop = new_env(env, false);
            exec(n->kids[0].get(), loop);
            // for (let i ...): every iteration has its own binding of the loop variables - a closure made in the body keeps the value of
            // ITS iteration (the copy is taken after the body and before the update expression, as JavaScript does)
            const bool per_iteration = n->kids[0]->k == K::VarDecl && n->kids[0]->s != "var";
            while (true) {
                if (n->kids[1]->k != K::Empty && !Interp::to_boolean(eval(n->kids[1].get(), loop))) break;
                Comp c = exec(n->kids[3].get(), loop);
                if (leaves_loop(c, mine)) { if (c == Comp::Break && jump_label.empty()) break; return c; }
                if (per_iteration) {
                    EnvP next = new_env(env, false);
                    next->vars = loop->vars;
                    next->is_const = loop->is_const;
                    loop = next;
                }
                if (n->kids[2]->k != K::Empty) eval(n->kids[2].get(), loop);
            }
            return Comp::Normal;
        }
        case K::ForIn: case K::ForOf: {
            const std::vector<std::string> mine = take_loop_labels();
            Value it = eval(n->kids[0].get(), env);
            std::vector<Value> items;
            if (it.type == Value::Obj) {
                if (n->k == K::ForOf) {
                    if (!Interp::iterate(it, items)) fail(n, "TypeError: for-of needs an array, a string, a Map or a Set");
                } else {
                    for (auto& k : Interp::own_keys(it)) items.push_back(Value(k));
                }
            }
            for (auto& item : items) {
                EnvP loop = new_env(env, false);
                if (n->flag && n->num != 0) loop->vars[n->s] = item;
                else set_var(n, loop, n->s, item);
                Comp c = exec(n->kids[1].get(), loop);
                if (leaves_loop(c, mine)) { if (c == Comp::Break && jump_label.empty()) break; return c; }
            }
            return Comp::Normal;
        }
        case K::While: {
            const std::vector<std::string> mine = take_loop_labels();
            while (Interp::to_boolean(eval(n->kids[0].get(), env))) {
                Comp c = exec(n->kids[1].get(), env);
                if (leaves_loop(c, mine)) { if (c == Comp::Break && jump_label.empty()) break; return c; }
            }
            return Comp::Normal;
        }
        case K::DoWhile: {
            const std::vector<std::string> mine = take_loop_labels();
            do {
                Comp c = exec(n->kids[0].get(), env);
                if (leaves_loop(c, mine)) { if (c == Comp::Break && jump_label.empty()) break; return c; }
            } while (Interp::to_boolean(eval(n->kids[1].get(), env)));
            return Comp::Normal;
        }
        case K::Break: jump_label = n->s; return Comp::Break;
        case K::Continue: jump_label = n->s; return Comp::Continue;
        case K::Labeled: {
            const Node* st = n->kids[0].get();
            std::vector<std::string> labels = {n->s};
            while (st->k == K::Labeled) { labels.push_back(st->s); st = st->kids[0].get(); }      // a: b: for (...)
            if (st->k == K::For || st->k == K::ForIn || st->k == K::ForOf || st->k == K::While || st->k == K::DoWhile) loop_labels = labels;    // a loop owns its labels (continue label)
            Comp c = exec(st, env);
            loop_labels.clear();
            if (c == Comp::Break && std::find(labels.begin(), labels.end(), jump_label) != labels.end()) { jump_label.clear(); return Comp::Normal; }
            return c;
        }
        case K::Throw: throw Thrown{eval(n->kids[0].get(), env)};
        case K::Try: {
            Comp c = Comp::Normal;
            auto run_finally = [&]() { if (n->kids.size() > 2) exec(n->kids[2].get(), env); };
            try {
                c = exec(n->kids[0].get(), env);
            } catch (Thrown& th) {
                if (n->kids[1]->k == K::Empty) { run_finally(); throw; }
                EnvP h = new_env(env, false);
                if (!n->s.empty()) h->vars[n->s] = th.v;
                c = exec(n->kids[1].get(), h);
            } catch (JsError& e) {
                if (n->kids[1]->k == K::Empty) { run_finally(); throw; }
                EnvP h = new_env(env, false);
                if (!n->s.empty()) h->vars[n->s] = in.make_error_object(e.what());
                c = exec(n->kids[1].get(), h);
            }
            run_finally();
            return c;
        }
        default:
            completion = eval(n, env);
            return Comp::Normal;
    }
}

// ------------------------------------------------------------------------------------------
// Interp
// ------------------------------------------------------------------------------------------
ObjP Interp::make_object() { return std::make_shared<Object>(); }
ObjP Interp::make_array(std::vector<Value> items) {
    auto o = std::make_shared<Object>();
    o->kind = Object::Array;
    o->arr = std::move(items);
    return o;
}
ObjP Interp::make_native(NativeFn fn) {
    auto o = std::make_shared<Object>();
    o->kind = Object::Native;
    o->native = std::move(fn);
    return o;
}
ObjP Interp::make_host(int tag, std::shared_ptr<void> payload) {
    auto o = std::make_shared<Object>();
    o->kind = Object::Host;
    o->host_tag = tag;
    o->host = std::move(payload);
    return o;
}

double Interp::to_number(const Value& v) {
    switch (v.type) {
        case Value::Undefined: return NAN;
        case Value::Null: return 0;
        case Value::Bool: return v.b ? 1 : 0;
        case Value::Number: return v.num;
        case Value::String: {
            if (v.str.empty()) return 0;
            char* end;
            double d = strtod(v.str.c_str(), &end);
            while (*end == ' ') end++;
            return *end ? NAN : d;
        }
        case Value::Obj:
            if (v.obj->kind == Object::Array && v.obj->arr.size() == 1) return to_number(v.obj->arr[0]);
            if (v.obj->kind == Object::Array && v.obj->arr.empty()) return 0;
            return NAN;
    }
    return NAN;
}
bool Interp::to_boolean(const Value& v) {
    switch (v.type) {
        case Value::Undefined: case Value::Null: return false;
        case Value::Bool: return v.b;
        case Value::Number: return v.num != 0 && !std::isnan(v.num);
        case Value::String: return !v.str.empty();
        case Value::Obj: return true;
    }
    return false;
}
std::string Interp::to_string(const Value& v) {
    switch (v.type) {
        case Value::Undefined: return "undefined";
        case Value::Null: return "null";
        case Value::Bool: return v.b ? "true" : "false";
        case Value::Number: {
            if (std::isnan(v.num)) return "NaN";
            if (std::isinf(v.num)) return v.num > 0 ? "Infinity" : "-Infinity";
            if (v.num == 0) return "0";
            // Number::toString: the shortest digit string that reads back as the same double, laid out as ECMAScript does
            // (plain decimals for 1e-7 <= |x| < 1e21, otherwise d.ddde+N)
            char buf[40];
            int prec = 1;
            for (; prec <= 17; prec++) {
                snprintf(buf, sizeof buf, "%.*e", prec - 1, std::fabs(v.num));
                if (strtod(buf, nullptr) == std::fabs(v.num)) break;
            }
            std::string digits;
            for (const char* c = buf; *c && *c != 'e'; c++) if (*c != '.') digits += *c;
            const int exp10 = atoi(strchr(buf, 'e') + 1);        // value = 0.d1d2... x 10^(exp10 + 1)
            const int n = exp10 + 1, k = (int)digits.size();
            std::string out = v.num < 0 ? "-" : "";
            if (k <= n && n <= 21) out += digits + std::string((size_t)(n - k), '0');
            else if (0 < n && n <= 21) out += digits.substr(0, (size_t)n) + "." + digits.substr((size_t)n);
            else if (-6 < n && n <= 0) out += "0." + std::string((size_t)(-n), '0') + digits;
            else {
                out += digits.substr(0, 1);
                if (k > 1) out += "." + digits.substr(1);
                out += std::string("e") + (n - 1 >= 0 ? "+" : "-") + std::to_string(std::abs(n - 1));
            }
            return out;
        }
        case Value::String: return v.str;
        case Value::Obj:
            if (v.obj->kind == Object::Array) {
                std::string out;
                for (size_t i = 0; i < v.obj->arr.size(); i++) {
                    if (i) out += ",";
                    if (v.obj->arr[i].type != Value::Undefined && v.obj->arr[i].type != Value::Null) out += to_string(v.obj->arr[i]);
                }
                return out;
            }
            if (v.obj->kind == Object::Function || v.obj->kind == Object::Native) return "function";
            return "[object Object]";
    }
    return "";
}

static Value math1(double (*f)(double), std::vector<Value>& a) { return Value(f(a.empty() ? NAN : Interp::to_number(a[0]))); }

static Value make_collection(Interp& ip, int kind, std::vector<Value>& args);

// ---- JSON -------------------------------------------------------------------------------------
namespace {
struct JsonReader {
    const std::string& s;
    size_t p = 0;
    [[noreturn]] void bad(const char* what) { throw JsError(std::string("SyntaxError: JSON.parse: ") + what + " at position " + std::to_string(p)); }
    void ws() { while (p < s.size() && (s[p] == ' ' || s[p] == '\t' || s[p] == '\n' || s[p] == '\r')) p++; }
    Value value() {
        ws();
        if (p >= s.size()) bad("unexpected end of input");
        const char c = s[p];
        if (c == '{') {
            p++;
            auto o = Interp::make_object();
            ws();
            if (p < s.size() && s[p] == '}') { p++; return Value(o); }
            for (;;) {
                ws();
                if (p >= s.size() || s[p] != '"') bad("expected a property name");
                const std::string key = string();
                ws();
                if (p >= s.size() || s[p] != ':') bad("expected ':'");
                p++;
                o->props[key] = value();
                ws();
                if (p < s.size() && s[p] == ',') { p++; continue; }
                if (p < s.size() && s[p] == '}') { p++; return Value(o); }
                bad("expected ',' or '}'");
            }
        }
        if (c == '[') {
            p++;
            std::vector<Value> items;
            ws();
            if (p < s.size() && s[p] == ']') { p++; return Value(Interp::make_array(items)); }
            for (;;) {
                items.push_back(value());
                ws();
                if (p < s.size() && s[p] == ',') { p++; continue; }
                if (p < s.size() && s[p] == ']') { p++; return Value(Interp::make_array(items)); }
                bad("expected ',' or ']'");
            }
        }
        if (c == '"') return Value(string());
        if (s.compare(p, 4, "true") == 0) { p += 4; return Value(true); }
        if (s.compare(p, 5, "false") == 0) { p += 5; return Value(false); }
        if (s.compare(p, 4, "null") == 0) { p += 4; return Value::null(); }
        if (c == '-' || isdigit((unsigned char)c)) {
            char* end = nullptr;
            const double v = strtod(s.c_str() + p, &end);
            if (end == s.c_str() + p) bad("bad number");
            p = (size_t)(end - s.c_str());
            return Value(v);
        }
        bad("unexpected token");
    }
    std::string string() {
        p++;                         // opening quote
        std::string out;
        while (p < s.size() && s[p] != '"') {
            if (s[p] == '\\' && p + 1 < s.size()) {
                p++;
                const char e = s[p];
                if (e == 'u' || e == 'n' || e == 't' || e == 'r' || e == 'b' || e == 'f') { int line = 0; read_escape(s, p, out, line); }
                else { out += e; p++; }          // \" \\ \/
            } else out += s[p++];
        }
        if (p >= s.size()) bad("unterminated string");
        p++;
        return out;
    }
};

void json_quote(const std::string& s, std::string& out) {
    out += '"';
    for (unsigned char c : s) {
        switch (c) {
            case '"': out += "\\\""; break;
            case '\\': out += "\\\\"; break;
            case '\n': out += "\\n"; break;
            case '\t': out += "\\t"; break;
            case '\r': out += "\\r"; break;
            case '\b': out += "\\b"; break;
            case '\f': out += "\\f"; break;
            default:
                if (c < 0x20) { char buf[8]; snprintf(buf, sizeof buf, "\\u%04x", c); out += buf; }
                else out += (char)c;
        }
    }
    out += '"';
}

// JSON.stringify(value [, null, indent]); undefined / functions are skipped in objects and become null in arrays
bool json_write(Interp& ip, const Value& v, const std::string& indent, const std::string& cur, std::string& out, int depth) {
    if (depth > 64) throw JsError("TypeError: cyclic object value");
    switch (v.type) {
        case Value::Undefined: return false;
        case Value::Null: out += "null"; return true;
        case Value::Bool: out += v.b ? "true" : "false"; return true;
        case Value::Number: out += std::isfinite(v.num) ? ip.to_string(v) : "null"; return true;
        case Value::String: json_quote(v.str, out); return true;
        case Value::Obj: break;
    }
    if (v.obj->kind == Object::Function || v.obj->kind == Object::Native) return false;
    const std::string inner = cur + indent;
    const char* nl = indent.empty() ? "" : "\n";
    if (v.obj->kind == Object::Array) {
        if (v.obj->arr.empty()) { out += "[]"; return true; }
        out += "[";
        for (size_t i = 0; i < v.obj->arr.size(); i++) {
            out += i ? "," : "";
            out += nl + inner;
            if (!json_write(ip, v.obj->arr[i], indent, inner, out, depth + 1)) out += "null";
        }
        out += nl + cur + "]";
        return true;
    }
    {
        auto tj = v.obj->props.find("toJSON");
        if (tj != v.obj->props.end() && tj->second.type == Value::Obj) return json_write(ip, ip.call(tj->second, v, {}), indent, cur, out, depth + 1);
    }
    std::string body;
    bool any = false;
    for (auto& k : Interp::own_keys(v)) {
        auto it = v.obj->props.find(k);
        Value item = it != v.obj->props.end() ? it->second : Value();
        std::string piece;
        if (!json_write(ip, item, indent, inner, piece, depth + 1)) continue;
        body += any ? "," : "";
        body += nl + inner;
        json_quote(k, body);
        body += indent.empty() ? ":" : ": ";
        body += piece;
        any = true;
    }
    out += any ? "{" + body + nl + cur + "}" : "{}";
    return true;
}
}  // namespace

Interp::Interp() {
    globals = std::make_shared<Env>();
    globals->function_scope = true;
    auto math = make_object();
#define M1(nm, fn) math->props[nm] = Value(make_native([](Interp&, const Value&, std::vector<Value>& a) { return math1(fn, a); }))
    M1("sin", std::sin); M1("cos", std::cos); M1("tan", std::tan); M1("asin", std::asin); M1("acos", std::acos);
    M1("atan", std::atan); M1("sqrt", std::sqrt); M1("cbrt", std::cbrt); M1("exp", std::exp); M1("log", std::log);
    M1("log2", std::log2); M1("log10", std::log10); M1("abs", std::fabs); M1("floor", std::floor); M1("ceil", std::ceil);
    M1("trunc", std::trunc); M1("sinh", std::sinh); M1("cosh", std::cosh); M1("tanh", std::tanh);
#undef M1
    math->props["round"] = Value(make_native([](Interp&, const Value&, std::vector<Value>& a) {
        const double x = a.empty() ? NAN : to_number(a[0]);
        if (!std::isfinite(x) || std::fabs(x) >= 4503599627370496.0) return Value(x);
        const double r = std::floor(x + 0.5);                       // ties go up
        return Value(r == 0 && (x < 0 || std::signbit(x)) ? -0.0 : r);
    }));
    math->props["fround"] = Value(make_native([](Interp&, const Value&, std::vector<Value>& a) { return Value((double)(float)(a.empty() ? NAN : to_number(a[0]))); }));
    math->props["expm1"] = Value(make_native([](Interp&, const Value&, std::vector<Value>& a) { return Value(std::expm1(a.empty() ? NAN : to_number(a[0]))); }));
    math->props["log1p"] = Value(make_native([](Interp&, const Value&, std::vector<Value>& a) { return Value(std::log1p(a.empty() ? NAN : to_number(a[0]))); }));
    math->props["asinh"] = Value(make_native([](Interp&, const Value&, std::vector<Value>& a) { return Value(std::asinh(a.empty() ? NAN : to_number(a[0]))); }));
    math->props["acosh"] = Value(make_native([](Interp&, const Value&, std::vector<Value>& a) { return Value(std::acosh(a.empty() ? NAN : to_number(a[0]))); }));
    math->props["atanh"] = Value(make_native([](Interp&, const Value&, std::vector<Value>& a) { return Value(std::atanh(a.empty() ? NAN : to_number(a[0]))); }));
    math->props["sign"] = Value(make_native([](Interp&, const Value&, std::vector<Value>& a) { double x = a.empty() ? NAN : to_number(a[0]); return Value(x > 0 ? 1.0 : (x < 0 ? -1.0 : x)); }));
    math->props["atan2"] = Value(make_native([](Interp&, const Value&, std::vector<Value>& a) { return Value(std::atan2(a.size() > 0 ? to_number(a[0]) : NAN, a.size() > 1 ? to_number(a[1]) : NAN)); }));
    math->props["pow"] = Value(make_native([](Interp&, const Value&, std::vector<Value>& a) { return Value(std::pow(a.size() > 0 ? to_number(a[0]) : NAN, a.size() > 1 ? to_number(a[1]) : NAN)); }));
    math->props["hypot"] = Value(make_native([](Interp&, const Value&, std::vector<Value>& a) { double s = 0; for (auto& v : a) { double x = to_number(v); s += x * x; } return Value(std::sqrt(s)); }));
    math->props["min"] = Value(make_native([](Interp&, const Value&, std::vector<Value>& a) { double m = INFINITY; for (auto& v : a) { double x = to_number(v); if (std::isnan(x)) return Value(NAN); m = std::min(m, x); } return Value(m); }));
    math->props["max"] = Value(make_native([](Interp&, const Value&, std::vector<Value>& a) { double m = -INFINITY; for (auto& v : a) { double x = to_number(v); if (std::isnan(x)) return Value(NAN); m = std::max(m, x); } return Value(m); }));
    math->props["PI"] = Value(3.14159265358979323846);
    math->props["E"] = Value(2.7182818284590452354);
    math->props["LN2"] = Value(0.69314718055994530942);
    math->props["LN10"] = Value(2.30258509299404568402);
    math->props["LOG2E"] = Value(1.4426950408889634074);
    math->props["LOG10E"] = Value(0.43429448190325182765);
    math->props["SQRT2"] = Value(1.41421356237309504880);
    math->props["SQRT1_2"] = Value(0.70710678118654752440);
    globals->vars["Math"] = Value(math);
    globals->vars["NaN"] = Value(NAN);
    globals->vars["Infinity"] = Value(INFINITY);
    globals->vars["isNaN"] = Value(make_native([](Interp&, const Value&, std::vector<Value>& a) { return Value(std::isnan(a.empty() ? NAN : to_number(a[0]))); }));
    globals->vars["isFinite"] = Value(make_native([](Interp&, const Value&, std::vector<Value>& a) { return Value((bool)std::isfinite(a.empty() ? NAN : to_number(a[0]))); }));
    auto parse_float = make_native([](Interp& ip, const Value&, std::vector<Value>& a) {
        if (a.empty()) return Value(NAN);
        const std::string s = ip.to_string(a[0]);
        size_t b = s.find_first_not_of(" \t\n\r");
        if (b == std::string::npos) return Value(NAN);
        if (s.compare(b, 2, "0x") == 0 || s.compare(b, 2, "0X") == 0) return Value(0.0);      // parseFloat reads the leading zero only
        char* end = nullptr;
        const double v = strtod(s.c_str() + b, &end);
        return Value(end == s.c_str() + b ? NAN : v);
    });
    auto parse_int = make_native([](Interp& ip, const Value&, std::vector<Value>& a) {
        if (a.empty()) return Value(NAN);
        std::string s = ip.to_string(a[0]);
        size_t b = s.find_first_not_of(" \t\n\r");
        if (b == std::string::npos) return Value(NAN);
        bool neg = false;
        if (s[b] == '+' || s[b] == '-') { neg = s[b] == '-'; b++; }
        int radix = a.size() > 1 && !a[1].is_undefined() ? (int)to_number(a[1]) : 0;
        if ((radix == 0 || radix == 16) && (s.compare(b, 2, "0x") == 0 || s.compare(b, 2, "0X") == 0)) { radix = 16; b += 2; }
        if (radix == 0) radix = 10;
        if (radix < 2 || radix > 36) return Value(NAN);
        double v = 0;
        bool any = false;
        for (; b < s.size(); b++) {
            const char c = s[b];
            const int d = c >= '0' && c <= '9' ? c - '0' : (c >= 'a' && c <= 'z' ? c - 'a' + 10 : (c >= 'A' && c <= 'Z' ? c - 'A' + 10 : 99));
            if (d >= radix) break;
            v = v * radix + d;
            any = true;
        }
        return Value(any ? (neg ? -v : v) : NAN);
    });
    globals->vars["parseFloat"] = Value(parse_float);
    globals->vars["parseInt"] = Value(parse_int);
    auto number_ctor = make_native([](Interp&, const Value&, std::vector<Value>& a) { return Value(a.empty() ? 0.0 : to_number(a[0])); });
    number_ctor->props["isFinite"] = Value(make_native([](Interp&, const Value&, std::vector<Value>& a) { return Value(!a.empty() && a[0].type == Value::Number && std::isfinite(a[0].num)); }));
    number_ctor->props["isNaN"] = Value(make_native([](Interp&, const Value&, std::vector<Value>& a) { return Value(!a.empty() && a[0].type == Value::Number && std::isnan(a[0].num)); }));
    number_ctor->props["isInteger"] = Value(make_native([](Interp&, const Value&, std::vector<Value>& a) { return Value(!a.empty() && a[0].type == Value::Number && std::isfinite(a[0].num) && a[0].num == std::floor(a[0].num)); }));
    number_ctor->props["isSafeInteger"] = Value(make_native([](Interp&, const Value&, std::vector<Value>& a) { return Value(!a.empty() && a[0].type == Value::Number && std::isfinite(a[0].num) && a[0].num == std::floor(a[0].num) && std::fabs(a[0].num) <= 9007199254740991.0); }));
    number_ctor->props["parseFloat"] = Value(parse_float);
    number_ctor->props["parseInt"] = Value(parse_int);
    number_ctor->props["EPSILON"] = Value(2.220446049250313e-16);
    number_ctor->props["MAX_SAFE_INTEGER"] = Value(9007199254740991.0);
    number_ctor->props["MIN_SAFE_INTEGER"] = Value(-9007199254740991.0);
    number_ctor->props["MAX_VALUE"] = Value(1.7976931348623157e308);
    number_ctor->props["MIN_VALUE"] = Value(5e-324);
    number_ctor->props["POSITIVE_INFINITY"] = Value(INFINITY);
    number_ctor->props["NEGATIVE_INFINITY"] = Value(-INFINITY);
    number_ctor->props["NaN"] = Value(NAN);
    globals->vars["Number"] = Value(number_ctor);
    globals->vars["String"] = Value(make_native([](Interp& ip, const Value&, std::vector<Value>& a) { return Value(a.empty() ? std::string() : ip.to_string(a[0])); }));
    // `${value}`: an object with its own toString() (then valueOf()) prints as that; number objects and everything else as in "" + value
    globals->vars["$$template_string"] = Value(make_native([](Interp& ip, const Value&, std::vector<Value>& a) {
        if (a.empty()) return Value(std::string("undefined"));
        if (a[0].type == Value::Obj && a[0].obj->kind == Object::Plain) {
            for (const char* name : {"toString", "valueOf"}) {
                Value fn;
                for (const Object* c = a[0].obj.get(); c && fn.is_undefined(); c = c->proto.get()) { auto it = c->props.find(name); if (it != c->props.end()) fn = it->second; }
                if (fn.type != Value::Obj || (fn.obj->kind != Object::Function && fn.obj->kind != Object::Native)) continue;
                Value r = ip.call(fn, a[0], {});
                if (r.type != Value::Obj) return Value(ip.to_string(r));
            }
        }
        return a[0].type == Value::Obj && a[0].obj->kind == Object::Host ? a[0] : Value(ip.to_string(a[0]));
    }));
    globals->vars["Boolean"] = Value(make_native([](Interp&, const Value&, std::vector<Value>& a) { return Value(!a.empty() && to_boolean(a[0])); }));
    auto array_ctor = make_native([](Interp&, const Value&, std::vector<Value>& a) {
        if (a.size() == 1 && a[0].type == Value::Number) {
            if (a[0].num < 0 || a[0].num > 1e7) throw JsError("RangeError: invalid array length");
            return Value(make_array(std::vector<Value>((size_t)a[0].num)));
        }
        return Value(make_array(a));
    });
    array_ctor->props["isArray"] = Value(make_native([](Interp&, const Value&, std::vector<Value>& a) { return Value(!a.empty() && a[0].type == Value::Obj && a[0].obj->kind == Object::Array); }));
    array_ctor->props["of"] = Value(make_native([](Interp&, const Value&, std::vector<Value>& a) { return Value(make_array(a)); }));
    // Array.from(array | string | {length: n} [, (v, i) => ...])
    array_ctor->props["from"] = Value(make_native([](Interp& ip, const Value&, std::vector<Value>& a) {
        std::vector<Value> items;
        if (!a.empty() && Interp::iterate(a[0], items)) {}                  // array, string, Map, Set
        else if (!a.empty() && a[0].type == Value::Obj) {
            auto len = a[0].obj->props.find("length");
            const double n = len == a[0].obj->props.end() ? 0.0 : to_number(len->second);
            if (n > 1e7) throw JsError("RangeError: invalid array length");
            for (size_t i = 0; n > 0 && i < (size_t)n; i++) {
                auto it = a[0].obj->props.find(std::to_string(i));
                items.push_back(it == a[0].obj->props.end() ? Value() : it->second);
            }
        }
        if (a.size() > 1 && !a[1].is_undefined())
            for (size_t i = 0; i < items.size(); i++) items[i] = ip.call(a[1], Value(), {items[i], Value((double)i)});
        return Value(make_array(items));
    }));
    // Array.prototype.slice.call(arguments, 1) and friends: every array method, applied to the array handed over as `this`
    auto array_proto = make_object();
    array_proto->get_hook = [](Interp&, const std::string& key, Value& out) {
        out = Value(make_native([key](Interp& ip, const Value& self, std::vector<Value>& args) {
            if (self.type != Value::Obj || self.obj->kind != Object::Array) throw JsError("TypeError: Array.prototype." + key + " needs an array as 'this'");
            Evaluator ev(ip, "<call>");
            Value m = ev.get_prop(nullptr, self, key);
            if (m.type != Value::Obj) throw JsError("TypeError: Array.prototype." + key + " is not a function");
            return ip.call(m, self, args);
        }));
        return true;
    };
    array_ctor->props["prototype"] = Value(array_proto);
    globals->vars["Array"] = Value(array_ctor);
    builtin_array = array_ctor;
    auto object_ctor = make_native([](Interp&, const Value&, std::vector<Value>&) { return Value(make_object()); });
    object_ctor->props["keys"] = Value(make_native([](Interp&, const Value&, std::vector<Value>& a) {
        std::vector<Value> keys;
        if (!a.empty()) for (auto& k : Interp::own_keys(a[0])) keys.push_back(Value(k));
        return Value(make_array(keys));
    }));
    auto own_keys = [](const Value& v) { return Interp::own_keys(v); };
    auto own_value = [](Interp& ip, const Value& v, const std::string& k) {
        size_t idx = 0;
        bool is_index = !k.empty() && k.find_first_not_of("0123456789") == std::string::npos;
        if (is_index) idx = (size_t)strtoull(k.c_str(), nullptr, 10);
        if (v.obj->kind == Object::Array && is_index) return idx < v.obj->arr.size() ? v.obj->arr[idx] : Value();
        auto it = v.obj->props.find(k);
        if (it != v.obj->props.end()) return it->second;
        auto g = v.obj->props.find("$get$" + k);                         // an accessor of the object literal: its value
        return g == v.obj->props.end() ? Value() : ip.call(g->second, v, {});
    };
    object_ctor->props["values"] = Value(make_native([own_keys, own_value](Interp& ip, const Value&, std::vector<Value>& a) {
        std::vector<Value> out;
        if (!a.empty()) for (auto& k : own_keys(a[0])) out.push_back(own_value(ip, a[0], k));
        return Value(make_array(out));
    }));
    object_ctor->props["entries"] = Value(make_native([own_keys, own_value](Interp& ip, const Value&, std::vector<Value>& a) {
        std::vector<Value> out;
        if (!a.empty()) for (auto& k : own_keys(a[0])) out.push_back(Value(make_array({Value(k), own_value(ip, a[0], k)})));
        return Value(make_array(out));
    }));
    object_ctor->props["fromEntries"] = Value(make_native([](Interp& ip, const Value&, std::vector<Value>& a) {
        auto o = make_object();
        if (!a.empty() && a[0].type == Value::Obj && a[0].obj->kind == Object::Array)
            for (auto& e : a[0].obj->arr)
                if (e.type == Value::Obj && e.obj->kind == Object::Array && e.obj->arr.size() >= 2) o->props[ip.to_string(e.obj->arr[0])] = e.obj->arr[1];
        return Value(o);
    }));
    object_ctor->props["assign"] = Value(make_native([own_keys, own_value](Interp& ip, const Value&, std::vector<Value>& a) {
        if (a.empty() || a[0].type != Value::Obj) throw JsError("TypeError: Object.assign needs an object");
        for (size_t i = 1; i < a.size(); i++)
            if (a[i].type == Value::Obj) for (auto& k : own_keys(a[i])) a[0].obj->props[k] = own_value(ip, a[i], k);
        return a[0];
    }));
    object_ctor->props["create"] = Value(make_native([](Interp&, const Value&, std::vector<Value>& a) {
        auto o = make_object();
        if (!a.empty() && a[0].type == Value::Obj) o->proto = a[0].obj;
        else if (a.empty() || a[0].type != Value::Null) throw JsError("TypeError: Object.create needs an object or null");
        return Value(o);
    }));
    object_ctor->props["getPrototypeOf"] = Value(make_native([](Interp&, const Value&, std::vector<Value>& a) {
        if (a.empty() || a[0].type != Value::Obj || !a[0].obj->proto) return Value::null();
        return Value(a[0].obj->proto);
    }));
    object_ctor->props["setPrototypeOf"] = Value(make_native([](Interp&, const Value&, std::vector<Value>& a) {
        if (a.size() < 2 || a[0].type != Value::Obj) throw JsError("TypeError: Object.setPrototypeOf needs an object");
        for (const Object* c = a[1].type == Value::Obj ? a[1].obj.get() : nullptr; c; c = c->proto.get())
            if (c == a[0].obj.get()) throw JsError("TypeError: cyclic prototype chain");
        a[0].obj->proto = a[1].type == Value::Obj ? a[1].obj : nullptr;
        return a[0];
    }));
    // Object.defineProperty(o, key, {value} | {get, set}): attributes (enumerable, writable, configurable) are not modelled
    object_ctor->props["defineProperty"] = Value(make_native([](Interp& ip, const Value&, std::vector<Value>& a) {
        if (a.size() < 3 || a[0].type != Value::Obj || a[2].type != Value::Obj) throw JsError("TypeError: Object.defineProperty needs an object, a key and a descriptor");
        const std::string key = ip.to_string(a[1]);
        auto& d = a[2].obj->props;
        if (d.count("get") || d.count("set")) {
            if (d.count("get") && d["get"].type == Value::Obj) a[0].obj->props["$get$" + key] = d["get"];
            if (d.count("set") && d["set"].type == Value::Obj) a[0].obj->props["$set$" + key] = d["set"];
        } else a[0].obj->props[key] = d.count("value") ? d["value"] : Value();
        if (d.count("enumerable") && to_boolean(d["enumerable"])) a[0].obj->hidden.erase(key); else a[0].obj->hidden.insert(key);
        return a[0];
    }));
    for (const char* same : {"freeze", "seal", "preventExtensions"})      // no property attributes in this interpreter: the object is handed back
        object_ctor->props[same] = Value(make_native([](Interp&, const Value&, std::vector<Value>& a) { return a.empty() ? Value() : a[0]; }));
    globals->vars["Object"] = Value(object_ctor);
    builtin_object = object_ctor;
    auto json = make_object();
    json->props["parse"] = Value(make_native([](Interp& ip, const Value&, std::vector<Value>& a) {
        const std::string text = a.empty() ? "undefined" : ip.to_string(a[0]);
        JsonReader r{text};
        Value v = r.value();
        r.ws();
        if (r.p != text.size()) r.bad("unexpected characters after the value");
        return v;
    }));
    json->props["stringify"] = Value(make_native([](Interp& ip, const Value&, std::vector<Value>& a) {
        std::string indent;
        if (a.size() > 2 && a[2].type == Value::Number) indent.assign((size_t)std::max(0.0, std::min(10.0, a[2].num)), ' ');
        else if (a.size() > 2 && a[2].type == Value::String) indent = a[2].str.substr(0, 10);
        std::string out;
        if (a.empty() || !json_write(ip, a[0], indent, "", out, 0)) return Value();
        return Value(out);
    }));
    globals->vars["JSON"] = Value(json);
    globals->vars["Map"] = Value(make_native([](Interp& ip, const Value&, std::vector<Value>& a) { return make_collection(ip, 1, a); }));
    globals->vars["Set"] = Value(make_native([](Interp& ip, const Value&, std::vector<Value>& a) { return make_collection(ip, 2, a); }));
    globals->vars["WeakMap"] = globals->vars["Map"];
    globals->vars["WeakSet"] = globals->vars["Set"];
    // Float64Array(n | array): a plain array of numbers (no fixed length, no buffer views)
    globals->vars["Float64Array"] = Value(make_native([](Interp&, const Value&, std::vector<Value>& a) {
        std::vector<Value> items;
        if (!a.empty() && a[0].type == Value::Number) {
            if (a[0].num < 0 || a[0].num > 1e7) throw JsError("RangeError: invalid typed array length");
            items.assign((size_t)a[0].num, Value(0.0));
        } else if (!a.empty() && Interp::iterate(a[0], items)) {
            for (auto& v : items) v = Value(to_number(v));
        }
        return Value(make_array(items));
    }));
    builtin_function = make_native([](Interp&, const Value&, std::vector<Value>&) -> Value { throw JsError("EvalError: the Function constructor is not available to metric scripts"); });
    globals->vars["Function"] = Value(builtin_function);
    // Error and its kinds: `new TypeError(msg)` -> {message, name}, `e instanceof Error`; errors the interpreter raises itself are
    // caught as objects of the matching kind (make_error_object)
    for (const char* kind : {"Error", "TypeError", "RangeError", "ReferenceError", "SyntaxError", "EvalError"}) {
        const std::string name = kind;
        auto ctor = std::make_shared<Object>();
        ctor->kind = Object::Native;
        std::weak_ptr<Object> self = ctor;
        ctor->native = [name, self](Interp& ip, const Value&, std::vector<Value>& a) {
            auto e = make_object();
            e->klass = self.lock();
            e->props["message"] = Value(a.empty() || a[0].is_undefined() ? std::string() : ip.to_string(a[0]));
            e->props["name"] = Value(name);
            return Value(e);
        };
        if (name != "Error") ctor->klass = error_types["Error"];
        error_types[name] = ctor;
        globals->vars[name] = Value(ctor);
    }
    auto console = make_object();
    auto logfn = make_native([](Interp& ip, const Value&, std::vector<Value>& a) {
        std::string line;
        for (size_t i = 0; i < a.size(); i++) { if (i) line += " "; line += ip.to_string(a[i]); }
        if (getenv("GRT_JS_CONSOLE")) fprintf(stderr, "[js] %s\n", line.c_str());
        return Value();
    });
    console->props["log"] = Value(logfn);
    console->props["warn"] = Value(logfn);
    console->props["error"] = Value(logfn);
    globals->vars["console"] = Value(console);
    globals->vars["print"] = Value(logfn);
    globals->vars["globalThis"] = Value(make_object());
}

Interp::~Interp() {
    // break closure <-> environment reference cycles
    std::set<Env*> seen;
    std::function<void(Env*)> clear = [&](Env* e) {
        if (!e || !seen.insert(e).second) return;
        auto vars = std::move(e->vars);
        e->vars.clear();
        for (auto& kv : vars)
            if (kv.second.type == Value::Obj && kv.second.obj->env) clear(kv.second.obj->env.get());
    };
    clear(globals.get());
}

// Enumerable own property names in JavaScript's order; an accessor of an object literal ("$get$v" / "$set$v") counts once, as "v"
std::vector<std::string> Interp::own_keys(const Value& v) {
    std::vector<std::string> keys;
    if (v.type != Value::Obj) return keys;
    if (v.obj->kind == Object::Array) for (size_t i = 0; i < v.obj->arr.size(); i++) keys.push_back(std::to_string(i));
    std::set<std::string> accessors;
    for (auto& kv : v.obj->props) {
        if (kv.first.compare(0, 5, "$get$") == 0 || kv.first.compare(0, 5, "$set$") == 0) {
            const std::string name = kv.first.substr(5);
            if (accessors.insert(name).second && !v.obj->props.count(name) && !v.obj->hidden.count(name)) keys.push_back(name);
        } else if (!v.obj->hidden.count(kv.first)) keys.push_back(kv.first);
    }
    return keys;
}

bool Interp::iterate(const Value& v, std::vector<Value>& out) {
    if (v.type == Value::String) { for (char c : v.str) out.push_back(Value(std::string(1, c))); return true; }
    if (v.type != Value::Obj) return false;
    if (v.obj->kind == Object::Array) { out = v.obj->arr; return true; }
    if (v.obj->is_collection == 1) { for (auto& e : v.obj->entries) out.push_back(Value(make_array({e.first, e.second}))); return true; }
    if (v.obj->is_collection == 2) { for (auto& e : v.obj->entries) out.push_back(e.first); return true; }
    return false;
}

// SameValueZero: the key comparison of Map and Set (NaN equals NaN, +0 equals -0, objects by identity)
static bool same_value_zero(const Value& a, const Value& b) {
    if (a.type != b.type) return false;
    switch (a.type) {
        case Value::Undefined: case Value::Null: return true;
        case Value::Bool: return a.b == b.b;
        case Value::Number: return a.num == b.num || (std::isnan(a.num) && std::isnan(b.num));
        case Value::String: return a.str == b.str;
        case Value::Obj: return a.obj == b.obj;
    }
    return false;
}

// new Map([[k, v], ...]) / new Set([v, ...]): methods as own natives of the object (get / set / has / delete / clear / add / forEach / keys /
// values / entries, `size` through a property hook); iteration through Interp::iterate
static Value make_collection(Interp& ip, int kind, std::vector<Value>& args) {
    auto o = Interp::make_object();
    o->is_collection = kind;
    std::weak_ptr<Object> weak = o;
    auto find = [](Object& c, const Value& k) { for (size_t i = 0; i < c.entries.size(); i++) if (same_value_zero(c.entries[i].first, k)) return (long)i; return -1L; };
    auto put = [find](Object& c, const Value& k, const Value& v) { long i = find(c, k); if (i >= 0) c.entries[(size_t)i].second = v; else c.entries.emplace_back(k, v); };
    o->get_hook = [weak](Interp&, const std::string& key, Value& out) {
        auto c = weak.lock();
        if (key == "size" && c) { out = Value((double)c->entries.size()); return true; }
        return false;
    };
    auto method = [&](const char* name, std::function<Value(Object&, Interp&, std::vector<Value>&)> fn) {
        o->props[name] = Value(Interp::make_native([weak, fn](Interp& ip2, const Value&, std::vector<Value>& a) { auto c = weak.lock(); if (!c) throw JsError("TypeError: collection is gone"); return fn(*c, ip2, a); }));
        o->hidden.insert(name);
    };
    method("has", [find](Object& c, Interp&, std::vector<Value>& a) { return Value(find(c, a.empty() ? Value() : a[0]) >= 0); });
    method("delete", [find](Object& c, Interp&, std::vector<Value>& a) { long i = find(c, a.empty() ? Value() : a[0]); if (i < 0) return Value(false); c.entries.erase(c.entries.begin() + i); return Value(true); });
    method("clear", [](Object& c, Interp&, std::vector<Value>&) { c.entries.clear(); return Value(); });
    method("forEach", [kind](Object& c, Interp& ip2, std::vector<Value>& a) {
        if (a.empty()) throw JsError("TypeError: callback required");
        auto snapshot = c.entries;
        for (auto& e : snapshot) ip2.call(a[0], Value(), {kind == 1 ? e.second : e.first, e.first});
        return Value();
    });
    method("keys", [](Object& c, Interp&, std::vector<Value>&) { std::vector<Value> out; for (auto& e : c.entries) out.push_back(e.first); return Value(Interp::make_array(out)); });
    method("values", [kind](Object& c, Interp&, std::vector<Value>&) { std::vector<Value> out; for (auto& e : c.entries) out.push_back(kind == 1 ? e.second : e.first); return Value(Interp::make_array(out)); });
    method("entries", [kind](Object& c, Interp&, std::vector<Value>&) { std::vector<Value> out; for (auto& e : c.entries) out.push_back(Value(Interp::make_array({e.first, kind == 1 ? e.second : e.first}))); return Value(Interp::make_array(out)); });
    if (kind == 1) {
        method("get", [find](Object& c, Interp&, std::vector<Value>& a) { long i = find(c, a.empty() ? Value() : a[0]); return i < 0 ? Value() : c.entries[(size_t)i].second; });
        o->props["set"] = Value(Interp::make_native([weak, put](Interp&, const Value&, std::vector<Value>& a) { auto c = weak.lock(); put(*c, a.empty() ? Value() : a[0], a.size() > 1 ? a[1] : Value()); return Value(c); }));
        o->hidden.insert("set");
    } else {
        o->props["add"] = Value(Interp::make_native([weak, put](Interp&, const Value&, std::vector<Value>& a) { auto c = weak.lock(); put(*c, a.empty() ? Value() : a[0], Value()); return Value(c); }));
        o->hidden.insert("add");
    }
    if (!args.empty() && !args[0].is_undefined() && args[0].type != Value::Null) {
        std::vector<Value> items;
        if (!Interp::iterate(args[0], items)) throw JsError("TypeError: the argument of a Map or Set constructor must be iterable");
        for (auto& it : items) {
            if (kind == 2) { put(*o, it, Value()); continue; }
            if (it.type != Value::Obj || it.obj->kind != Object::Array) throw JsError("TypeError: a Map is built from [key, value] pairs");
            put(*o, it.obj->arr.size() > 0 ? it.obj->arr[0] : Value(), it.obj->arr.size() > 1 ? it.obj->arr[1] : Value());
        }
    }
    (void)ip;
    return Value(o);
}

Value Interp::make_error_object(const std::string& what) {
    auto e = make_object();
    std::string kind = "Error";
    for (auto& kv : error_types) {
        const size_t at = what.find(kv.first + ": ");
        if (kv.first != "Error" && at != std::string::npos) { kind = kv.first; break; }
    }
    e->klass = error_types[kind];
    e->props["message"] = Value(what);
    e->props["name"] = Value(kind);
    return Value(e);
}

void Interp::set_global(const std::string& name, const Value& v) { globals->vars[name] = v; }
Value Interp::get_global(const std::string& name) {
    auto it = globals->vars.find(name);
    return it == globals->vars.end() ? Value() : it->second;
}
void Interp::set_global_getter(const std::string& name, std::function<Value(Interp&)> getter) {
    global_getters[name] = std::move(getter);
}

Value Interp::eval(const std::string& source, const std::string& name) {
    Lexer lx(source, name);
    lx.run();
    Parser ps(lx.toks, name);
    programs.push_back(ps.parse_program());
    Program* prog = programs.back().get();
    Evaluator ev(*this, name);
    steps = 0;
    try {
        ev.hoist(prog->top, globals);
        Comp c = ev.exec_block(prog->body, globals);
        (void)c;
    } catch (Thrown& th) {
        std::string msg = th.v.type == Value::Obj && th.v.obj->props.count("message") ? to_string(th.v.obj->props["message"]) : to_string(th.v);
        throw JsError(name + ": uncaught exception: " + msg);
    }
    return ev.completion;
}

Value Interp::call(const Value& fn, const Value& self, std::vector<Value> args) {
    // errors raised while the function runs name the script that defines it, like the errors of its first evaluation
    const bool scripted = fn.type == Value::Obj && fn.obj->kind == Object::Function && fn.obj->def && !fn.obj->def->source.empty();
    Evaluator ev(*this, scripted ? fn.obj->def->source : std::string("<call>"));
    try {
        return ev.call_function(nullptr, fn, self, args);
    } catch (Thrown& th) {
        std::string msg = th.v.type == Value::Obj && th.v.obj->props.count("message") ? to_string(th.v.obj->props["message"]) : to_string(th.v);
        throw JsError("uncaught exception: " + msg);
    }
}

}  // namespace minijs
