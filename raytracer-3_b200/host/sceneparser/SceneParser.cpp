/* sceneparser/SceneParser.cpp — see SceneParser.hpp. Hand-written tokenizer and recursive-descent parser. */
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <fstream>
#include <map>
#include <sstream>

#include <unistd.h>

#include <CppDebugger.hpp>

#include "entities/Object.hpp"
#include "entities/Sphere.hpp"
#include "entities/Triangle.hpp"
#include "SceneParser.hpp"

using namespace RayTracer;
using namespace CppDebugger::SeverityValues;

namespace {
    /* ---- values ---- */
    enum class Type { boolean, integer, uinteger, real, vec3 };
    struct Value {
        Type type = Type::real;
        double s = 0.0;            /* scalar payload (exact for 32-bit ints) */
        float v[3] = { 0, 0, 0 };
        static Value scalar(Type t, double x) { Value r; r.type = t; r.s = x; return r; }
        static Value vector(float x, float y, float z) { Value r; r.type = Type::vec3; r.v[0] = x; r.v[1] = y; r.v[2] = z; return r; }
        float f() const { return (float) s; }
    };

    /* ---- tokens ---- */
    enum class Tok { end, id, number, string, punct, format, at, raw };
    struct Token { Tok kind = Tok::end; std::string text; bool is_float = false; size_t line = 1; };

    struct Source {
        std::string name, text, dir;
        size_t pos = 0, line = 1;
    };

    /* Inline data lives in temporary files that the object entities re-read when they are pre-rendered
     * (reference Object.cpp:136-178 opens the file again), so the files go when the process does. */
    void keep_until_exit(const std::string& path) {
        static std::vector<std::string>* files = nullptr;
        if (!files) {
            files = new std::vector<std::string>();
            std::atexit([] { for (const std::string& f : *files) { unlink(f.c_str()); } });
        }
        files->push_back(path);
    }

    class Parser {
        std::vector<Source> stack;          /* #include stack; back() is being read */
        Token look;                         /* one token of lookahead */
        bool has_look = false;
        ParsedScene& out;
        std::map<std::string, std::string> data_paths;                   /* data id -> .obj file */
        std::map<std::string, std::map<std::string, Value>> fields;      /* entity id | "global" -> key -> value */

        [[noreturn]] void error(const std::string& msg, size_t line = 0) {
            const Source& s = stack.back();
            DLOG(fatal, s.name + ":" + std::to_string(line ? line : s.line) + ": " + msg);
            throw CppDebugger::Fatal(msg); /* not reached: DLOG(fatal) throws */
        }

        /* -- lexer -- */
        void skip_space() {
            for (;;) {
                Source& s = stack.back();
                while (s.pos < s.text.size() && (s.text[s.pos] == ' ' || s.text[s.pos] == '\t' || s.text[s.pos] == '\r' || s.text[s.pos] == '\n')) {
                    if (s.text[s.pos] == '\n') { s.line++; }
                    s.pos++;
                }
                if (s.text.compare(s.pos, 2, "/*") == 0) {
                    size_t e = s.text.find("*/", s.pos + 2);
                    if (e == std::string::npos) { error("unterminated comment"); }
                    for (size_t i = s.pos; i < e; i++) { if (s.text[i] == '\n') { s.line++; } }
                    s.pos = e + 2;
                    continue;
                }
                if (s.text.compare(s.pos, 2, "//") == 0) {
                    while (s.pos < s.text.size() && s.text[s.pos] != '\n') { s.pos++; }
                    continue;
                }
                if (s.pos >= s.text.size() && stack.size() > 1) { stack.pop_back(); continue; } /* end of an included file */
                return;
            }
        }

        std::string read_string(char quote) {
            Source& s = stack.back();
            std::string r;
            s.pos++;
            while (s.pos < s.text.size() && s.text[s.pos] != quote) {
                char c = s.text[s.pos++];
                if (c == '\n') { s.line++; }
                if (c == '\\' && s.pos < s.text.size()) {
                    char e = s.text[s.pos++];
                    c = e == 'n' ? '\n' : e == 'r' ? '\r' : e == 't' ? '\t' : e;
                }
                r += c;
            }
            if (s.pos >= s.text.size()) { error("unterminated string"); }
            s.pos++;
            return r;
        }

        Token lex() {
            skip_space();
            Source& s = stack.back();
            Token t;
            t.line = s.line;
            if (s.pos >= s.text.size()) { return t; }
            const char c = s.text[s.pos];
            auto is_id_start = [](char ch) { return (ch >= 'a' && ch <= 'z') || (ch >= 'A' && ch <= 'Z') || ch == '_'; };
            auto is_digit = [](char ch) { return ch >= '0' && ch <= '9'; };
            if (is_id_start(c)) {
                size_t e = s.pos;
                /* identifiers may contain dashes (SceneLang.md:15); a dash must be followed by an identifier character */
                while (e < s.text.size() && (is_id_start(s.text[e]) || is_digit(s.text[e]) ||
                                             (s.text[e] == '-' && e + 1 < s.text.size() && (is_id_start(s.text[e + 1]) || is_digit(s.text[e + 1]))))) { e++; }
                t.kind = Tok::id; t.text = s.text.substr(s.pos, e - s.pos); s.pos = e;
                return t;
            }
            if (is_digit(c) || (c == '.' && s.pos + 1 < s.text.size() && is_digit(s.text[s.pos + 1]))) {
                size_t e = s.pos;
                while (e < s.text.size() && is_digit(s.text[e])) { e++; }
                if (e < s.text.size() && s.text[e] == '.') { t.is_float = true; e++; while (e < s.text.size() && is_digit(s.text[e])) { e++; } }
                if (e < s.text.size() && (s.text[e] == 'e' || s.text[e] == 'E')) {
                    size_t x = e + 1;
                    if (x < s.text.size() && (s.text[x] == '-' || s.text[x] == '+')) { x++; }
                    if (x < s.text.size() && is_digit(s.text[x])) { t.is_float = true; e = x; while (e < s.text.size() && is_digit(s.text[e])) { e++; } }
                }
                t.kind = Tok::number; t.text = s.text.substr(s.pos, e - s.pos); s.pos = e;
                return t;
            }
            if (c == '.' && s.pos + 1 < s.text.size() && is_id_start(s.text[s.pos + 1])) {
                /* either a data format (.obj) or the second half of a reference (entity.key); the parser decides from context */
                size_t e = s.pos + 1;
                while (e < s.text.size() && (is_id_start(s.text[e]) || is_digit(s.text[e]))) { e++; }
                t.kind = Tok::format; t.text = s.text.substr(s.pos + 1, e - s.pos - 1); s.pos = e;
                return t;
            }
            if (c == '"' || c == '\'') { t.kind = Tok::string; t.text = read_string(c); return t; }
            if (c == '@' || c == '#') {
                size_t e = s.pos + 1;
                while (e < s.text.size() && is_id_start(s.text[e])) { e++; }
                t.kind = Tok::at; t.text = s.text.substr(s.pos, e - s.pos); s.pos = e;
                return t;
            }
            static const char* two[] = { "==", "!=", ">=", "<=", "&&", "||", "<<", ">>" };
            for (const char* op : two) { if (s.text.compare(s.pos, 2, op) == 0) { t.kind = Tok::punct; t.text = op; s.pos += 2; return t; } }
            t.kind = Tok::punct; t.text = std::string(1, c); s.pos++;
            return t;
        }

        const Token& peek() { if (!has_look) { look = lex(); has_look = true; } return look; }
        Token next() { Token t = peek(); has_look = false; return t; }
        bool is_punct(const char* p) { return peek().kind == Tok::punct && peek().text == p; }
        bool accept(const char* p) { if (is_punct(p)) { next(); return true; } return false; }
        void expect(const char* p) { if (!accept(p)) { error(std::string("expected '") + p + "', got '" + peek().text + "'", peek().line); } }
        std::string expect_id(const char* what) {
            if (peek().kind != Tok::id) { error(std::string("expected ") + what + ", got '" + peek().text + "'", peek().line); }
            return next().text;
        }

        /* Raw text of an inline data block: everything up to the matching unescaped '}' (SceneLang.md:51). */
        std::string read_raw_block() {
            if (has_look) { error("internal: lookahead before a raw block"); }
            Source& s = stack.back();
            std::string r;
            while (s.pos < s.text.size() && s.text[s.pos] != '}') {
                char c = s.text[s.pos++];
                if (c == '\n') { s.line++; }
                if (c == '\\' && s.pos < s.text.size() && (s.text[s.pos] == '\\' || s.text[s.pos] == '{' || s.text[s.pos] == '}')) { c = s.text[s.pos++]; }
                r += c;
            }
            if (s.pos >= s.text.size()) { error("unterminated data block"); }
            s.pos++;
            return r;
        }

        std::string resolve(const std::string& path) {
            if (!path.empty() && (path[0] == '/' || (path.size() > 1 && path[1] == ':'))) { return path; }
            return stack.back().dir.empty() ? path : stack.back().dir + "/" + path;
        }

        /* -- @statements and #include -- */
        bool directive() {
            if (peek().kind != Tok::at) { return false; }
            Token d = next();
            if (d.text == "#include") {
                if (peek().kind != Tok::string) { error("#include needs a string", d.line); }
                const std::string path = resolve(next().text);
                std::ifstream in(path, std::ios::binary);
                if (!in.is_open()) { error("cannot open included file '" + path + "'", d.line); }
                std::stringstream ss; ss << in.rdbuf();
                if (stack.size() > 32) { error("#include nested too deeply", d.line); }
                Source src; src.name = path; src.text = ss.str();
                size_t slash = path.find_last_of('/');
                src.dir = slash == std::string::npos ? "" : path.substr(0, slash);
                has_look = false;
                stack.push_back(src);
                return true;
            }
            if (peek().kind != Tok::id && peek().kind != Tok::string) { error(d.text + " needs an identifier or a string", d.line); }
            const Token arg = next();
            if (d.text == "@error") { error(arg.kind == Tok::string ? arg.text : "error '" + arg.text + "' raised by the scene file", d.line); }
            else if (d.text == "@warning") { out.warnings.push_back(stack.back().name + ":" + std::to_string(d.line) + ": " + arg.text); }
            else if (d.text != "@ignore" && d.text != "@suppress") { error("unknown statement '" + d.text + "'", d.line); }
            return true;
        }
        void directives() { while (directive()) {} }

        /* -- expressions (C precedence) -- */
        static bool type_name(const std::string& s, Type* t) {
            if (s == "bool") { *t = Type::boolean; } else if (s == "int") { *t = Type::integer; } else if (s == "uint") { *t = Type::uinteger; }
            else if (s == "float") { *t = Type::real; } else if (s == "vec3") { *t = Type::vec3; } else { return false; }
            return true;
        }
        Value cast(const Value& v, Type t, size_t line) {
            if (v.type == t) { return v; }
            if (t == Type::vec3) { return Value::vector(v.f(), v.f(), v.f()); }              /* a scalar fills the vector */
            if (v.type == Type::vec3) { error("a vec3 cannot be cast to a scalar type", line); }
            switch (t) {
                case Type::boolean: return Value::scalar(t, v.s != 0.0 ? 1.0 : 0.0);
                case Type::integer: return Value::scalar(t, (double) (int32_t) v.s);
                case Type::uinteger: return Value::scalar(t, (double) (uint32_t) (int64_t) v.s);
                default: return Value::scalar(t, (double) (float) v.s);
            }
        }
        static Type wider(Type a, Type b) {
            if (a == Type::vec3 || b == Type::vec3) { return Type::vec3; }
            if (a == Type::real || b == Type::real) { return Type::real; }
            if (a == Type::uinteger || b == Type::uinteger) { return Type::uinteger; }
            if (a == Type::integer || b == Type::integer) { return Type::integer; }
            return Type::boolean;
        }
        Value arith(const std::string& op, const Value& a, const Value& b, size_t line) {
            const Type t = wider(a.type, b.type);
            if (t == Type::vec3) {
                if (op != "+" && op != "-" && op != "*" && op != "/") { error("operator '" + op + "' is not defined for vec3", line); }
                const Value x = cast(a, Type::vec3, line), y = cast(b, Type::vec3, line);
                Value r = x;
                for (int k = 0; k < 3; k++) {
                    r.v[k] = op == "+" ? x.v[k] + y.v[k] : op == "-" ? x.v[k] - y.v[k] : op == "*" ? x.v[k] * y.v[k] : x.v[k] / y.v[k];
                }
                return r;
            }
            if (t == Type::real) {
                const float x = a.f(), y = b.f();   /* all SceneLang types are 32-bit (SceneLang.md:141) */
                if (op == "+") { return Value::scalar(t, x + y); } if (op == "-") { return Value::scalar(t, x - y); }
                if (op == "*") { return Value::scalar(t, x * y); } if (op == "/") { return Value::scalar(t, x / y); }
                if (op == "%") { return Value::scalar(t, std::fmod(x, y)); }
                error("operator '" + op + "' needs integer operands", line);
            }
            const int64_t x = (int64_t) a.s, y = (int64_t) b.s;
            int64_t r = 0;
            if (op == "+") { r = x + y; } else if (op == "-") { r = x - y; } else if (op == "*") { r = x * y; }
            else if (op == "/" || op == "%") { if (y == 0) { error("division by zero", line); } r = op == "/" ? x / y : x % y; }
            else if (op == "&") { r = x & y; } else if (op == "|") { r = x | y; } else if (op == "^") { r = x ^ y; }
            else if (op == "<<") { r = x << (y & 31); } else if (op == ">>") { r = x >> (y & 31); }
            const Type rt = t == Type::boolean ? Type::integer : t;
            return cast(Value::scalar(Type::real, (double) r), rt, line);
        }
        Value compare(const std::string& op, const Value& a, const Value& b, size_t line) {
            if (a.type == Type::vec3 || b.type == Type::vec3) {
                if (op != "==" && op != "!=") { error("vec3 values can only be compared with == and !=", line); }
                const Value x = cast(a, Type::vec3, line), y = cast(b, Type::vec3, line);
                const bool eq = x.v[0] == y.v[0] && x.v[1] == y.v[1] && x.v[2] == y.v[2];
                return Value::scalar(Type::boolean, (op == "==") == eq ? 1.0 : 0.0);
            }
            const double x = a.s, y = b.s;
            const bool r = op == "==" ? x == y : op == "!=" ? x != y : op == "<" ? x < y : op == ">" ? x > y : op == "<=" ? x <= y : x >= y;
            return Value::scalar(Type::boolean, r ? 1.0 : 0.0);
        }

        Value primary() {
            const Token t = next();
            if (t.kind == Tok::number) {
                if (t.is_float) { return Value::scalar(Type::real, (double) std::strtof(t.text.c_str(), nullptr)); }
                return Value::scalar(Type::integer, (double) std::strtoll(t.text.c_str(), nullptr, 10));
            }
            if (t.kind == Tok::punct && t.text == "(") {
                Type ct;
                if (peek().kind == Tok::id && type_name(peek().text, &ct)) {           /* (type) expr */
                    next(); expect(")");
                    return cast(unary(), ct, t.line);
                }
                Value v;
                if (!triple(")", v)) { v = expression(); }
                expect(")");
                return v;
            }
            if (t.kind == Tok::id) {
                if (t.text == "true") { return Value::scalar(Type::boolean, 1.0); }
                if (t.text == "false") { return Value::scalar(Type::boolean, 0.0); }
                if (is_punct("(")) { return call(t); }
                if (peek().kind == Tok::format) {                                       /* entity.key or global.key */
                    const std::string key = next().text;
                    auto e = fields.find(t.text);
                    if (e == fields.end()) { error("reference to unknown entity '" + t.text + "'", t.line); }
                    auto f = e->second.find(key);
                    if (f == e->second.end()) { error("'" + t.text + "' has no parameter '" + key + "' (yet)", t.line); }
                    return f->second;
                }
                error("unexpected identifier '" + t.text + "' in an expression (references are <entity>.<key>)", t.line);
            }
            error("unexpected '" + t.text + "' in an expression", t.line);
        }
        Value call(const Token& name) {
            expect("(");
            std::vector<Value> args;
            if (!is_punct(")")) { do { args.push_back(expression()); } while (accept(",")); }
            expect(")");
            auto need = [&](size_t n) { if (args.size() != n) { error(name.text + "() takes " + std::to_string(n) + " argument(s)", name.line); } };
            auto real = [&](size_t i) { return cast(args[i], Type::real, name.line).f(); };
            if (name.text == "vec3") { need(3); return Value::vector(real(0), real(1), real(2)); }
            if (name.text == "sqrt") { need(1); return Value::scalar(Type::real, std::sqrt(real(0))); }
            if (name.text == "sin") { need(1); return Value::scalar(Type::real, std::sin(real(0))); }
            if (name.text == "cos") { need(1); return Value::scalar(Type::real, std::cos(real(0))); }
            if (name.text == "abs") { need(1); return Value::scalar(Type::real, std::fabs(real(0))); }
            if (name.text == "min") { need(2); return Value::scalar(Type::real, std::fmin(real(0), real(1))); }
            if (name.text == "max") { need(2); return Value::scalar(Type::real, std::fmax(real(0), real(1))); }
            if (name.text == "tan") { need(1); return Value::scalar(Type::real, std::tan(real(0))); }
            if (name.text == "pow") { need(2); return Value::scalar(Type::real, std::pow(real(0), real(1))); }
            if (name.text == "floor") { need(1); return Value::scalar(Type::real, std::floor(real(0))); }
            if (name.text == "ceil") { need(1); return Value::scalar(Type::real, std::ceil(real(0))); }
            if (name.text == "radians") { need(1); return Value::scalar(Type::real, real(0) * (float) (M_PI / 180.0)); }
            /* vector helpers (the specification names "vector operations and related build-in functions", its appendix C is empty) */
            auto vec = [&](size_t i) -> const float* {
                if (args[i].type != Type::vec3) { error(name.text + "() needs vec3 arguments", name.line); }
                return args[i].v;
            };
            if (name.text == "dot") { need(2); const float *a = vec(0), *b = vec(1); return Value::scalar(Type::real, (a[0] * b[0] + a[1] * b[1]) + a[2] * b[2]); }
            if (name.text == "length") { need(1); const float* a = vec(0); return Value::scalar(Type::real, std::sqrt((a[0] * a[0] + a[1] * a[1]) + a[2] * a[2])); }
            if (name.text == "cross") {
                need(2);
                const float *a = vec(0), *b = vec(1);
                return Value::vector(a[1] * b[2] - b[1] * a[2], a[2] * b[0] - b[2] * a[0], a[0] * b[1] - b[0] * a[1]);
            }
            if (name.text == "normalize") {
                need(1);
                const float* a = vec(0);
                const float inv = 1.0f / std::sqrt((a[0] * a[0] + a[1] * a[1]) + a[2] * a[2]);
                return Value::vector(a[0] * inv, a[1] * inv, a[2] * inv);
            }
            error("unknown function '" + name.text + "'", name.line);
        }
        Value unary() {
            if (peek().kind == Tok::punct) {
                const Token t = peek();
                if (t.text == "-" || t.text == "+" || t.text == "!" || t.text == "~") {
                    next();
                    const Value v = unary();
                    if (t.text == "+") { return v; }
                    if (t.text == "-") {
                        if (v.type == Type::vec3) { return Value::vector(-v.v[0], -v.v[1], -v.v[2]); }
                        return Value::scalar(v.type == Type::boolean || v.type == Type::uinteger ? Type::integer : v.type, -v.s);
                    }
                    if (v.type == Type::vec3 || v.type == Type::real) { error("operator '" + t.text + "' needs an integer or bool operand", t.line); }
                    if (t.text == "!") { return Value::scalar(Type::boolean, v.s == 0.0 ? 1.0 : 0.0); }
                    return cast(Value::scalar(Type::real, (double) ~(int64_t) v.s), v.type == Type::boolean ? Type::integer : v.type, t.line);
                }
            }
            return primary();
        }
        /* binary operators by rising precedence level */
        Value binary(int level) {
            static const std::vector<std::vector<std::string>> levels = {
                { "||" }, { "&&" }, { "|" }, { "^" }, { "&" }, { "==", "!=" }, { "<", ">", "<=", ">=" }, { "<<", ">>" }, { "+", "-" }, { "*", "/", "%" } };
            if (level == (int) levels.size()) { return unary(); }
            Value left = binary(level + 1);
            for (;;) {
                if (peek().kind != Tok::punct) { return left; }
                const Token op = peek();
                bool mine = false;
                for (const std::string& o : levels[level]) { if (o == op.text) { mine = true; } }
                if (!mine) { return left; }
                next();
                const Value right = binary(level + 1);
                if (op.text == "||") { left = Value::scalar(Type::boolean, (cast(left, Type::boolean, op.line).s != 0 || cast(right, Type::boolean, op.line).s != 0) ? 1.0 : 0.0); }
                else if (op.text == "&&") { left = Value::scalar(Type::boolean, (cast(left, Type::boolean, op.line).s != 0 && cast(right, Type::boolean, op.line).s != 0) ? 1.0 : 0.0); }
                else if (level == 5 || level == 6) { left = compare(op.text, left, right, op.line); }
                else { left = arith(op.text, left, right, op.line); }
            }
        }
        Value expression() { return binary(0); }

        /* Three juxtaposed unary-level values followed by `term` make a vec3 (SceneLang.md:119); otherwise nothing is consumed. */
        bool triple(const char* term, Value& result) {
            const size_t saved_depth = stack.size(), saved_pos = stack.back().pos, saved_line = stack.back().line;
            const Token saved_look = look;
            const bool saved_has = has_look;
            const size_t saved_warnings = out.warnings.size();
            bool ok = true;
            Value c[3];
            try {
                for (int k = 0; k < 3 && ok; k++) {
                    if (is_punct(term)) { ok = false; break; }
                    c[k] = unary();
                    if (c[k].type == Type::vec3) { ok = false; }
                }
                if (ok && !is_punct(term)) { ok = false; }
            } catch (CppDebugger::Fatal&) { ok = false; }
            if (ok) { result = Value::vector(c[0].f(), c[1].f(), c[2].f()); return true; }
            if (stack.size() != saved_depth) { error("a value may not run across the end of an included file"); }
            stack.back().pos = saved_pos; stack.back().line = saved_line;
            look = saved_look; has_look = saved_has; out.warnings.resize(saved_warnings);
            return false;
        }

        /* A parameter value up to ';'. */
        Value value() {
            Value v;
            if (triple(";", v)) { return v; }
            return expression();
        }

        /* -- statements -- */
        void parameter(std::map<std::string, Value>& into, std::string* data_ref) {
            directives();
            Type declared = Type::real;
            bool has_type = false;
            std::string key = expect_id("a parameter name");
            const size_t line = peek().line;
            if (key == "data" && (is_punct(":") || peek().kind == Tok::id)) {              /* data [<key>]: .obj <id>; */
                if (peek().kind == Tok::id) { next(); }
                expect(":");
                if (peek().kind != Tok::format || peek().text != "obj") { error("expected a data format (.obj)", line); }
                next();
                const std::string id = expect_id("a data identifier");
                expect(";");
                if (!data_ref) { error("only entities can reference data", line); }
                if (!data_paths.count(id)) { error("reference to undefined data '" + id + "'", line); }
                *data_ref = id;
                return;
            }
            if (type_name(key, &declared) && peek().kind == Tok::id) { has_type = true; key = next().text; }
            expect(":");
            Value v = value();
            expect(";");
            if (has_type) {
                if (declared != Type::vec3 && v.type == Type::vec3) { error("parameter '" + key + "' is declared scalar but given a vec3", line); }
                v = cast(v, declared, line);
            }
            if (into.count(key)) { error("parameter '" + key + "' is defined twice", line); }
            into[key] = v;
        }

        const Value& required(const std::map<std::string, Value>& p, const std::string& entity, const char* key, size_t line) {
            auto it = p.find(key);
            if (it == p.end()) { error("entity '" + entity + "' is missing parameter '" + key + "'", line); }
            return it->second;
        }
        glm::vec3 vec(const Value& v, size_t line) { const Value x = cast(v, Type::vec3, line); return glm::vec3(x.v[0], x.v[1], x.v[2]); }
        uint32_t count(const Value& v, const char* key, size_t line) {
            if (v.type == Type::vec3 || v.s < 0 || v.s != std::floor(v.s) || v.s > 4294967295.0) { error(std::string("parameter '") + key + "' must be a non-negative integer", line); }
            return (uint32_t) v.s;
        }

        void entity_statement() {
            directives();
            if (is_punct("}")) { return; }
            const size_t line = peek().line;
            const std::string type = expect_id("an entity type");
            if (type != "triangle" && type != "sphere" && type != "object") { error("unknown entity type '" + type + "'", line); }
            const std::string id = expect_id("an entity identifier");
            if (id == "global") { error("'global' cannot be used as an entity identifier", line); }
            if (fields.count(id)) { error("entity '" + id + "' is defined twice", line); }
            expect("{");
            std::map<std::string, Value>& p = fields[id];
            std::string data_ref;
            while (!is_punct("}")) {
                if (peek().kind == Tok::end) { error("unterminated entity '" + id + "'", line); }
                parameter(p, &data_ref);
            }
            expect("}");
            ECS::RenderEntity* e = nullptr;
            if (type == "triangle") {
                e = (ECS::RenderEntity*) ECS::create_triangle(vec(required(p, id, "p1", line), line), vec(required(p, id, "p2", line), line),
                                                              vec(required(p, id, "p3", line), line), vec(required(p, id, "color", line), line));
            } else if (type == "sphere") {
                const uint32_t m = count(required(p, id, "n_meridians", line), "n_meridians", line), n = count(required(p, id, "n_parallels", line), "n_parallels", line);
                if (m < 1 || n < 3) { error("sphere '" + id + "' needs n_meridians >= 1 and n_parallels >= 3", line); }
                e = (ECS::RenderEntity*) ECS::create_sphere(vec(required(p, id, "center", line), line), cast(required(p, id, "radius", line), Type::real, line).f(), m, n,
                                                            vec(required(p, id, "color", line), line));
            } else {
                if (data_ref.empty()) { error("object '" + id + "' is missing its 'data' parameter", line); }
                e = (ECS::RenderEntity*) ECS::create_object(data_paths[data_ref], vec(required(p, id, "center", line), line),
                                                            cast(required(p, id, "scale", line), Type::real, line).f(), vec(required(p, id, "color", line), line));
            }
            out.entities.push_back(e);
            out.names.push_back(id);
        }

        void data_statement() {
            directives();
            if (is_punct("}")) { return; }
            const size_t line = peek().line;
            bool external = false;
            if (peek().kind == Tok::id && peek().text == "extern") { next(); external = true; }
            if (peek().kind != Tok::format || peek().text != "obj") { error("expected a data format (.obj), got '" + peek().text + "'", line); }
            next();
            const std::string id = expect_id("a data identifier");
            if (data_paths.count(id)) { error("data '" + id + "' is defined twice", line); }
            if (external) {
                expect(":");
                if (peek().kind != Tok::string) { error("extern data needs a path string", line); }
                data_paths[id] = resolve(next().text);
                expect(";");
                return;
            }
            if (!is_punct("{")) { error("expected '{' after the data identifier", line); }
            has_look = false;                                   /* the lexer stopped right behind '{' */
            const std::string text = read_raw_block();
            /* ECS::create_object reads a file (reference Object.cpp:81-122): inline data goes through a temporary one */
            char name[] = "/tmp/rt3_scene_XXXXXX.obj";
            const int fd = mkstemps(name, 4);
            if (fd < 0) { error("cannot create a temporary file for inline data '" + id + "'", line); }
            std::FILE* f = fdopen(fd, "wb");
            /* one record per line, without the block's indentation (the reference's loader is strict) */
            std::istringstream lines(text);
            std::string l;
            while (std::getline(lines, l)) {
                size_t b = l.find_first_not_of(" \t\r"), e2 = l.find_last_not_of(" \t\r");
                if (b == std::string::npos) { continue; }
                std::fprintf(f, "%s\n", l.substr(b, e2 - b + 1).c_str());
            }
            std::fclose(f);
            keep_until_exit(name);
            data_paths[id] = name;
        }

    public:
        explicit Parser(ParsedScene& o) : out(o) {}

        void run(const std::string& text, const std::string& dir, const std::string& name) {
            Source s; s.name = name; s.text = text; s.dir = dir;
            stack.push_back(s);
            try {
                for (;;) {
                    directives();
                    if (peek().kind == Tok::end) { break; }
                    const size_t line = peek().line;
                    const std::string section = expect_id("a section name (data, entities or global)");
                    if (section != "data" && section != "entities" && section != "global") { error("unknown section '" + section + "'", line); }
                    expect("{");
                    while (!is_punct("}")) {
                        if (peek().kind == Tok::end) { error("unterminated section '" + section + "'", line); }
                        if (section == "data") { data_statement(); }
                        else if (section == "entities") { entity_statement(); }
                        else { parameter(fields["global"], nullptr); }
                    }
                    expect("}");
                }
            } catch (...) {
                /* entities created so far are released: the caller gets all or nothing */
                for (size_t i = 0; i < out.entities.size(); i++) {
                    if (out.entities[i]->type == ECS::et_object) { delete (ECS::Object*) out.entities[i]; }
                    else if (out.entities[i]->type == ECS::et_sphere) { delete (ECS::Sphere*) out.entities[i]; }
                    else { delete (ECS::Triangle*) out.entities[i]; }
                }
                out.entities.clear(); out.names.clear();
                throw;
            }
        }
    };
}

void SceneParser::parse_string(const std::string& text, const std::string& base_dir, ParsedScene& out, const std::string& name) {
    Parser p(out);
    p.run(text, base_dir, name);
}

void SceneParser::parse_file(const std::string& path, ParsedScene& out) {
    std::ifstream in(path, std::ios::binary);
    if (!in.is_open()) { DLOG(fatal, "Could not open scene file '" + path + "'"); }
    std::stringstream ss;
    ss << in.rdbuf();
    const size_t slash = path.find_last_of('/');
    parse_string(ss.str(), slash == std::string::npos ? "" : path.substr(0, slash), out, path);
}
