// interp.cpp — host interpreter for LuisaCompute ASTs; see interp.h.  TEST INFRASTRUCTURE ONLY.
#include "interp.h"

#include <cmath>
#include <cstring>
#include <stdexcept>
#include <string>
#include <unordered_map>

#include <luisa/ast/constant_data.h>
#include <luisa/ast/expression.h>
#include <luisa/ast/function_builder.h>
#include <luisa/ast/statement.h>
#include <luisa/ast/type.h>
#include <luisa/core/stl/format.h>

namespace refinterp {

using namespace luisa;
using namespace luisa::compute;
using Tag = Type::Tag;

namespace {

[[noreturn]] void fail(const std::string &what) { throw std::runtime_error("refinterp: " + what); }

/* ---- scalar lanes ------------------------------------------------------------------------------------------- */

struct Lane {
    Tag k{Tag::FLOAT32};
    union {
        bool b;
        int32_t i;
        uint32_t u;
        int64_t l;
        uint64_t ul;
        float f;
        double d;
    };
    Lane() : ul{0u} {}
};

size_t scalar_size(Tag k) {
    switch (k) {
        case Tag::BOOL: return 1u;
        case Tag::INT16:
        case Tag::UINT16:
        case Tag::FLOAT16: return 2u;
        case Tag::INT32:
        case Tag::UINT32:
        case Tag::FLOAT32: return 4u;
        case Tag::INT64:
        case Tag::UINT64:
        case Tag::FLOAT64: return 8u;
        default: fail("not a scalar tag");
    }
}

Lane mk(bool v) { Lane x; x.k = Tag::BOOL; x.b = v; return x; }
Lane mk(int32_t v) { Lane x; x.k = Tag::INT32; x.i = v; return x; }
Lane mk(uint32_t v) { Lane x; x.k = Tag::UINT32; x.u = v; return x; }
Lane mk(int64_t v) { Lane x; x.k = Tag::INT64; x.l = v; return x; }
Lane mk(uint64_t v) { Lane x; x.k = Tag::UINT64; x.ul = v; return x; }
Lane mk(float v) { Lane x; x.k = Tag::FLOAT32; x.f = v; return x; }
Lane mk(double v) { Lane x; x.k = Tag::FLOAT64; x.d = v; return x; }

template<typename F>
auto visit(const Lane &x, F &&f) {
    switch (x.k) {
        case Tag::BOOL: return f(x.b);
        case Tag::INT32: return f(x.i);
        case Tag::UINT32: return f(x.u);
        case Tag::INT64: return f(x.l);
        case Tag::UINT64: return f(x.ul);
        case Tag::FLOAT32: return f(x.f);
        case Tag::FLOAT64: return f(x.d);
        default: fail("unsupported scalar type (16-bit)");
    }
}

Lane conv(const Lane &x, Tag k) {
    if (x.k == k) { return x; }
    return visit(x, [k](auto v) -> Lane {
        switch (k) {
            case Tag::BOOL: return mk(static_cast<bool>(v));
            case Tag::INT32: return mk(static_cast<int32_t>(v));
            case Tag::UINT32:
                if constexpr (std::is_floating_point_v<decltype(v)>) {
                    return mk(static_cast<uint32_t>(static_cast<int64_t>(v)));
                } else {
                    return mk(static_cast<uint32_t>(v));
                }
            case Tag::INT64: return mk(static_cast<int64_t>(v));
            case Tag::UINT64: return mk(static_cast<uint64_t>(v));
            case Tag::FLOAT32: return mk(static_cast<float>(v));
            case Tag::FLOAT64: return mk(static_cast<double>(v));
            default: fail("unsupported conversion target");
        }
    });
}

Lane load(const std::byte *p, Tag k) {
    Lane x;
    x.k = k;
    switch (k) {
        case Tag::BOOL: x.b = *reinterpret_cast<const bool *>(p); break;
        case Tag::INT32:
        case Tag::UINT32:
        case Tag::FLOAT32: std::memcpy(&x.u, p, 4u); break;
        case Tag::INT64:
        case Tag::UINT64:
        case Tag::FLOAT64: std::memcpy(&x.ul, p, 8u); break;
        default: fail("unsupported scalar load");
    }
    return x;
}

void store(std::byte *p, const Lane &x) {
    switch (x.k) {
        case Tag::BOOL: *reinterpret_cast<bool *>(p) = x.b; break;
        case Tag::INT32:
        case Tag::UINT32:
        case Tag::FLOAT32: std::memcpy(p, &x.u, 4u); break;
        case Tag::INT64:
        case Tag::UINT64:
        case Tag::FLOAT64: std::memcpy(p, &x.ul, 8u); break;
        default: fail("unsupported scalar store");
    }
}

/* ---- values ------------------------------------------------------------------------------------------------- */

struct Val {
    const Type *t{nullptr};
    std::vector<std::byte> m;
    Val() = default;
    explicit Val(const Type *type) : t{type}, m(type == nullptr ? 0u : type->size(), std::byte{0}) {}
    std::byte *p() { return m.data(); }
    const std::byte *p() const { return m.data(); }
};

struct Ptr {
    std::byte *p{nullptr};
    const Type *t{nullptr};
};

/* element layout of scalars / vectors / matrices */
struct Shape {
    Tag elem{};
    uint32_t n{1u};   /* lanes of a scalar (1) or vector (2..4); matrices: N */
    bool matrix{false};
};

Shape shape_of(const Type *t) {
    if (t->is_scalar()) { return {t->tag(), 1u, false}; }
    if (t->is_vector()) { return {t->element()->tag(), t->dimension(), false}; }
    if (t->is_matrix()) { return {Tag::FLOAT32, t->dimension(), true}; }
    fail("expected a scalar, vector or matrix, got " + std::string{t->description()});
}

std::vector<Lane> lanes(const Val &v) {
    auto s = shape_of(v.t);
    std::vector<Lane> out;
    if (!s.matrix) {
        auto es = scalar_size(s.elem);
        for (auto i = 0u; i < s.n; i++) { out.push_back(load(v.p() + i * es, s.elem)); }
    } else {
        auto stride = (s.n == 3u ? 4u : s.n) * 4u;
        for (auto c = 0u; c < s.n; c++) {
            for (auto r = 0u; r < s.n; r++) { out.push_back(load(v.p() + c * stride + r * 4u, Tag::FLOAT32)); }
        }
    }
    return out;
}

Val from_lanes(const Type *t, const std::vector<Lane> &ls) {
    Val v{t};
    auto s = shape_of(t);
    if (!s.matrix) {
        if (ls.size() != s.n) { fail("lane count mismatch for " + std::string{t->description()}); }
        auto es = scalar_size(s.elem);
        for (auto i = 0u; i < s.n; i++) { store(v.p() + i * es, conv(ls[i], s.elem)); }
    } else {
        if (ls.size() != s.n * s.n) { fail("lane count mismatch for matrix"); }
        auto stride = (s.n == 3u ? 4u : s.n) * 4u;
        for (auto c = 0u; c < s.n; c++) {
            for (auto r = 0u; r < s.n; r++) { store(v.p() + c * stride + r * 4u, conv(ls[c * s.n + r], Tag::FLOAT32)); }
        }
    }
    return v;
}

size_t align_up(size_t x, size_t a) { return (x + a - 1u) / a * a; }

size_t member_offset(const Type *st, uint32_t index) {
    auto off = size_t{0u};
    auto members = st->members();
    for (auto i = 0u; i <= index; i++) {
        off = align_up(off, members[i]->alignment());
        if (i == index) { return off; }
        off += members[i]->size();
    }
    return off;
}

/* ---- promotion & arithmetic --------------------------------------------------------------------------------- */

int rank(Tag k) {
    switch (k) {
        case Tag::BOOL: return 0;
        case Tag::INT16: return 1;
        case Tag::UINT16: return 2;
        case Tag::INT32: return 3;
        case Tag::UINT32: return 4;
        case Tag::INT64: return 5;
        case Tag::UINT64: return 6;
        case Tag::FLOAT16: return 7;
        case Tag::FLOAT32: return 8;
        case Tag::FLOAT64: return 9;
        default: fail("rank of non-scalar");
    }
}

Tag promote(Tag a, Tag b) {
    auto k = rank(a) >= rank(b) ? a : b;
    return rank(k) < rank(Tag::INT32) ? Tag::INT32 : k;
}

template<typename T, bool integral = std::is_integral_v<T>>
struct unsigned_of { using type = T; };
template<typename T>
struct unsigned_of<T, true> { using type = std::make_unsigned_t<T>; };

template<typename T>
Lane arith(BinaryOp op, T a, T b) {
    using U = typename unsigned_of<T>::type;
    switch (op) {
        case BinaryOp::ADD:
            if constexpr (std::is_integral_v<T>) { return mk(static_cast<T>(static_cast<U>(a) + static_cast<U>(b))); }
            else { return mk(static_cast<T>(a + b)); }
        case BinaryOp::SUB:
            if constexpr (std::is_integral_v<T>) { return mk(static_cast<T>(static_cast<U>(a) - static_cast<U>(b))); }
            else { return mk(static_cast<T>(a - b)); }
        case BinaryOp::MUL:
            if constexpr (std::is_integral_v<T>) { return mk(static_cast<T>(static_cast<U>(a) * static_cast<U>(b))); }
            else { return mk(static_cast<T>(a * b)); }
        case BinaryOp::DIV:
            if constexpr (std::is_integral_v<T>) { return mk(b == T{0} ? T{0} : static_cast<T>(a / b)); }
            else { return mk(static_cast<T>(a / b)); }
        case BinaryOp::MOD:
            if constexpr (std::is_integral_v<T>) { return mk(b == T{0} ? T{0} : static_cast<T>(a % b)); }
            else { return mk(static_cast<T>(std::fmod(a, b))); }
        case BinaryOp::BIT_AND:
            if constexpr (std::is_integral_v<T>) { return mk(static_cast<T>(a & b)); } else { fail("& on float"); }
        case BinaryOp::BIT_OR:
            if constexpr (std::is_integral_v<T>) { return mk(static_cast<T>(a | b)); } else { fail("| on float"); }
        case BinaryOp::BIT_XOR:
            if constexpr (std::is_integral_v<T>) { return mk(static_cast<T>(a ^ b)); } else { fail("^ on float"); }
        case BinaryOp::SHL:
            if constexpr (std::is_integral_v<T>) { return mk(static_cast<T>(static_cast<U>(a) << (static_cast<U>(b) & (sizeof(T) * 8u - 1u)))); }
            else { fail("<< on float"); }
        case BinaryOp::SHR:
            if constexpr (std::is_integral_v<T>) { return mk(static_cast<T>(a >> (static_cast<U>(b) & (sizeof(T) * 8u - 1u)))); }
            else { fail(">> on float"); }
        case BinaryOp::AND: return mk(static_cast<bool>(a) && static_cast<bool>(b));
        case BinaryOp::OR: return mk(static_cast<bool>(a) || static_cast<bool>(b));
        case BinaryOp::LESS: return mk(a < b);
        case BinaryOp::GREATER: return mk(a > b);
        case BinaryOp::LESS_EQUAL: return mk(a <= b);
        case BinaryOp::GREATER_EQUAL: return mk(a >= b);
        case BinaryOp::EQUAL: return mk(a == b);
        case BinaryOp::NOT_EQUAL: return mk(a != b);
    }
    fail("unknown binary op");
}

Lane binary_lane(BinaryOp op, const Lane &a, const Lane &b) {
    auto ct = promote(a.k, b.k);
    auto ca = conv(a, ct);
    auto cb = conv(b, ct);
    switch (ct) {
        case Tag::INT32: return arith<int32_t>(op, ca.i, cb.i);
        case Tag::UINT32: return arith<uint32_t>(op, ca.u, cb.u);
        case Tag::INT64: return arith<int64_t>(op, ca.l, cb.l);
        case Tag::UINT64: return arith<uint64_t>(op, ca.ul, cb.ul);
        case Tag::FLOAT32: return arith<float>(op, ca.f, cb.f);
        case Tag::FLOAT64: return arith<double>(op, ca.d, cb.d);
        default: fail("unsupported arithmetic type");
    }
}

bool is_relational(BinaryOp op) { return op >= BinaryOp::LESS; }

/* column-major float matrices as lanes[c * n + r] */
Val binary(BinaryOp op, const Val &a, const Val &b, const Type *rt) {
    auto sa = shape_of(a.t);
    auto sb = shape_of(b.t);
    auto la = lanes(a);
    auto lb = lanes(b);
    if (sa.matrix || sb.matrix) {
        auto n = sa.matrix ? sa.n : sb.n;
        std::vector<Lane> out;
        if (sa.matrix && sb.matrix && op == BinaryOp::MUL) {// (A * B)[c] = A * B[c]
            for (auto c = 0u; c < n; c++) {
                for (auto r = 0u; r < n; r++) {
                    auto acc = 0.f;// lc: m[0] * v.x + m[1] * v.y + ...
                    for (auto k = 0u; k < n; k++) {
                        auto term = la[k * n + r].f * lb[c * n + k].f;
                        acc = k == 0u ? term : acc + term;
                    }
                    out.push_back(mk(acc));
                }
            }
            return from_lanes(rt, out);
        }
        if (sa.matrix && !sb.matrix && sb.n == n && op == BinaryOp::MUL) {// M * v = v.x * m[0] + v.y * m[1] + ...
            for (auto r = 0u; r < n; r++) {
                auto acc = 0.f;
                for (auto k = 0u; k < n; k++) {
                    auto term = lb[k].f * la[k * n + r].f;
                    acc = k == 0u ? term : acc + term;
                }
                out.push_back(mk(acc));
            }
            return from_lanes(rt, out);
        }
        if (sa.matrix && sb.matrix) {// component-wise +, -
            for (auto i = 0u; i < n * n; i++) { out.push_back(binary_lane(op, la[i], lb[i])); }
            return from_lanes(rt, out);
        }
        if (sa.matrix && sb.n == 1u) {
            for (auto i = 0u; i < n * n; i++) { out.push_back(binary_lane(op, la[i], lb[0])); }
            return from_lanes(rt, out);
        }
        if (sb.matrix && sa.n == 1u) {
            for (auto i = 0u; i < n * n; i++) { out.push_back(binary_lane(op, la[0], lb[i])); }
            return from_lanes(rt, out);
        }
        fail("unsupported matrix binary operation");
    }
    auto n = std::max(sa.n, sb.n);
    if ((sa.n != 1u && sa.n != n) || (sb.n != 1u && sb.n != n)) { fail("vector size mismatch in binary op"); }
    std::vector<Lane> out;
    for (auto i = 0u; i < n; i++) { out.push_back(binary_lane(op, la[sa.n == 1u ? 0u : i], lb[sb.n == 1u ? 0u : i])); }
    (void)is_relational;
    return from_lanes(rt, out);
}

Val unary(UnaryOp op, const Val &a, const Type *rt) {
    auto la = lanes(a);
    std::vector<Lane> out;
    for (auto &x : la) {
        switch (op) {
            case UnaryOp::PLUS: out.push_back(x); break;
            case UnaryOp::MINUS:
                out.push_back(visit(x, [](auto v) -> Lane {
                    using T = decltype(v);
                    if constexpr (std::is_same_v<T, bool>) { return mk(-static_cast<int32_t>(v)); }
                    else if constexpr (std::is_integral_v<T>) { return mk(static_cast<T>(std::make_unsigned_t<T>{0} - static_cast<std::make_unsigned_t<T>>(v))); }
                    else { return mk(static_cast<T>(-v)); }
                }));
                break;
            case UnaryOp::NOT: out.push_back(mk(!conv(x, Tag::BOOL).b)); break;
            case UnaryOp::BIT_NOT:
                out.push_back(visit(x, [](auto v) -> Lane {
                    using T = decltype(v);
                    if constexpr (std::is_same_v<T, bool>) { return mk(!v); }
                    else if constexpr (std::is_integral_v<T>) { return mk(static_cast<T>(~v)); }
                    else { fail("~ on float"); }
                }));
                break;
        }
    }
    return from_lanes(rt, out);
}

/* ---- builtin math: the reference's CUDA backend header is the specification ---------------------------------- */

float powi_impl(float x, int y) {// cuda_device_math.h:21-32
    auto r = 1.0f;
    auto neg = y < 0;
    auto ya = neg ? -y : y;
    while (ya != 0) {
        if (ya & 1) { r *= x; }
        x *= x;
        ya >>= 1;
    }
    return neg ? 1.0f / r : r;
}
float powf_impl(float x, float y) {// :33-36
    auto yi = static_cast<int>(y);
    return static_cast<float>(yi) == y ? powi_impl(x, yi) : std::pow(x, y);
}
bool isinf_impl(float x) { uint32_t u; std::memcpy(&u, &x, 4u); return u == 0x7f800000u || u == 0xff800000u; }
bool isnan_impl(float x) { uint32_t u; std::memcpy(&u, &x, 4u); return (u & 0x7f800000u) == 0x7f800000u && (u & 0x7fffffu) != 0u; }

}// namespace

/* ---- interpreter -------------------------------------------------------------------------------------------- */

namespace {

struct Slot {
    std::byte *ptr{nullptr};
    std::vector<std::byte> own;
    BufferArg buffer;
};

enum struct Flow { NORMAL, BREAK, CONTINUE, RETURN };

class Machine {

private:
    Function _f;
    std::unordered_map<uint32_t, Slot> _slots;
    Val _ret;

public:
    explicit Machine(Function f) : _f{f} {}
    Slot &slot(Variable v) {
        auto it = _slots.find(v.uid());
        if (it == _slots.end()) {
            Slot s;
            if (!v.is_resource()) {
                s.own.assign(std::max<size_t>(v.type()->size(), 1u), std::byte{0});// locals are zero-initialised
            }
            it = _slots.emplace(v.uid(), std::move(s)).first;
            if (!v.is_resource()) { it->second.ptr = it->second.own.data(); }
        }
        return it->second;
    }
    void bind_value(Variable v, const std::byte *data) {
        auto &s = slot(v);
        std::memcpy(s.ptr, data, v.type()->size());
    }
    void bind_reference(Variable v, std::byte *target) {
        Slot s;
        s.ptr = target;
        _slots[v.uid()] = std::move(s);
    }
    void bind_buffer(Variable v, BufferArg b) {
        Slot s;
        s.buffer = b;
        _slots[v.uid()] = std::move(s);
    }
    Val run() {
        exec(_f.body());
        return std::move(_ret);
    }

private:
    bool try_lvalue(const Expression *e, Ptr &out) {
        switch (e->tag()) {
            case Expression::Tag::REF: {
                auto v = static_cast<const RefExpr *>(e)->variable();
                if (v.is_resource()) { return false; }
                if (v.tag() != Variable::Tag::LOCAL && v.tag() != Variable::Tag::REFERENCE && v.tag() != Variable::Tag::SHARED) {
                    fail("builtin variables (thread ids) are not available in callables");
                }
                out = {slot(v).ptr, v.type()};
                return true;
            }
            case Expression::Tag::MEMBER: {
                auto m = static_cast<const MemberExpr *>(e);
                Ptr self;
                if (!try_lvalue(m->self(), self)) { return false; }
                if (m->is_swizzle()) {
                    if (m->swizzle_size() != 1u) { return false; }
                    auto es = scalar_size(self.t->element()->tag());
                    out = {self.p + m->swizzle_index(0u) * es, e->type()};
                    return true;
                }
                out = {self.p + member_offset(self.t, m->member_index()), e->type()};
                return true;
            }
            case Expression::Tag::ACCESS: {
                auto a = static_cast<const AccessExpr *>(e);
                if (a->range()->type()->is_buffer()) { return false; }
                Ptr range;
                if (!try_lvalue(a->range(), range)) { return false; }
                auto idx = static_cast<size_t>(conv(lanes(eval(a->index()))[0], Tag::UINT32).u);
                out = {range.p + element_offset(range.t, idx), e->type()};
                return true;
            }
            default: return false;
        }
    }

    static size_t element_offset(const Type *t, size_t idx) {
        if (t->is_vector()) {
            if (idx >= t->dimension()) { fail("vector index out of range"); }
            return idx * scalar_size(t->element()->tag());
        }
        if (t->is_matrix()) {
            if (idx >= t->dimension()) { fail("matrix column out of range"); }
            return idx * (t->dimension() == 3u ? 4u : t->dimension()) * 4u;
        }
        if (t->is_array()) {
            if (idx >= t->dimension()) { fail("array index out of range"); }
            return idx * align_up(t->element()->size(), t->element()->alignment());
        }
        fail("access into " + std::string{t->description()});
    }

    Val eval(const Expression *e) {
        Ptr lv;
        if (try_lvalue(e, lv)) {
            Val v{e->type()};
            std::memcpy(v.p(), lv.p, v.m.size());
            return v;
        }
        switch (e->tag()) {
            case Expression::Tag::UNARY: {
                auto u = static_cast<const UnaryExpr *>(e);
                return unary(u->op(), eval(u->operand()), e->type());
            }
            case Expression::Tag::BINARY: {
                auto b = static_cast<const BinaryExpr *>(e);
                return binary(b->op(), eval(b->lhs()), eval(b->rhs()), e->type());
            }
            case Expression::Tag::MEMBER: {
                auto m = static_cast<const MemberExpr *>(e);
                auto self = eval(m->self());
                Val out{e->type()};
                if (m->is_swizzle()) {
                    auto es = scalar_size(self.t->element()->tag());
                    for (auto i = 0u; i < m->swizzle_size(); i++) {
                        std::memcpy(out.p() + i * es, self.p() + m->swizzle_index(i) * es, es);
                    }
                } else {
                    std::memcpy(out.p(), self.p() + member_offset(self.t, m->member_index()), out.m.size());
                }
                return out;
            }
            case Expression::Tag::ACCESS: {
                auto a = static_cast<const AccessExpr *>(e);
                auto idx = static_cast<size_t>(conv(lanes(eval(a->index()))[0], Tag::UINT32).u);
                if (a->range()->type()->is_buffer()) { return buffer_read(a->range(), idx, e->type()); }
                auto range = eval(a->range());
                Val out{e->type()};
                std::memcpy(out.p(), range.p() + element_offset(range.t, idx), out.m.size());
                return out;
            }
            case Expression::Tag::LITERAL: {
                auto l = static_cast<const LiteralExpr *>(e);
                Val out{e->type()};
                luisa::visit(
                    [&out](auto v) {
                        if (sizeof(v) != out.m.size()) { fail("literal size mismatch"); }
                        std::memcpy(out.p(), &v, sizeof(v));
                    },
                    l->value().to_variant());
                return out;
            }
            case Expression::Tag::CONSTANT: {
                auto c = static_cast<const ConstantExpr *>(e);
                Val out{e->type()};
                std::memcpy(out.p(), c->data().raw(), out.m.size());
                return out;
            }
            case Expression::Tag::CAST: {
                auto c = static_cast<const CastExpr *>(e);
                auto src = eval(c->expression());
                if (c->op() == CastOp::BITWISE) {
                    Val out{e->type()};
                    if (out.m.size() != src.m.size()) { fail("bitwise cast between different sizes"); }
                    std::memcpy(out.p(), src.p(), out.m.size());
                    return out;
                }
                return from_lanes(e->type(), lanes(src));
            }
            case Expression::Tag::CALL: return call(static_cast<const CallExpr *>(e));
            default: fail("unsupported expression tag " + std::to_string(static_cast<int>(e->tag())));
        }
    }

    BufferArg buffer_of(const Expression *e) {
        if (e->tag() != Expression::Tag::REF) { fail("buffer expression is not a variable"); }
        auto v = static_cast<const RefExpr *>(e)->variable();
        auto it = _slots.find(v.uid());
        if (it == _slots.end() || it->second.buffer.data == nullptr) { fail("unbound buffer argument"); }
        return it->second.buffer;
    }

    Val buffer_read(const Expression *buffer, size_t idx, const Type *elem) {
        auto b = buffer_of(buffer);
        if (idx >= b.count) { fail("buffer read out of range"); }
        Val out{elem};
        std::memcpy(out.p(), b.data + idx * align_up(elem->size(), elem->alignment()), out.m.size());
        return out;
    }

    /* ---- calls ---- */
    template<typename F>
    Val map_f(const CallExpr *e, F &&fn) {// float -> float, per lane
        auto a = lanes(eval(e->arguments()[0]));
        std::vector<Lane> out;
        for (auto &x : a) { out.push_back(mk(static_cast<float>(fn(conv(x, Tag::FLOAT32).f)))); }
        return from_lanes(e->type(), out);
    }
    template<typename F>
    Val map_ff(const CallExpr *e, F &&fn) {
        auto a = lanes(eval(e->arguments()[0]));
        auto b = lanes(eval(e->arguments()[1]));
        auto n = std::max(a.size(), b.size());
        std::vector<Lane> out;
        for (auto i = 0u; i < n; i++) {
            out.push_back(mk(static_cast<float>(fn(conv(a[a.size() == 1u ? 0u : i], Tag::FLOAT32).f, conv(b[b.size() == 1u ? 0u : i], Tag::FLOAT32).f))));
        }
        return from_lanes(e->type(), out);
    }
    template<typename F>
    Val map_fff(const CallExpr *e, F &&fn) {
        auto a = lanes(eval(e->arguments()[0]));
        auto b = lanes(eval(e->arguments()[1]));
        auto c = lanes(eval(e->arguments()[2]));
        auto n = std::max({a.size(), b.size(), c.size()});
        std::vector<Lane> out;
        auto at = [](const std::vector<Lane> &v, size_t i) { return conv(v[v.size() == 1u ? 0u : i], Tag::FLOAT32).f; };
        for (auto i = 0u; i < n; i++) { out.push_back(mk(static_cast<float>(fn(at(a, i), at(b, i), at(c, i))))); }
        return from_lanes(e->type(), out);
    }
    /* min / max / abs / clamp keep the operand type */
    static Lane lane_min(const Lane &a, const Lane &b) {
        auto ct = promote(a.k, b.k);
        auto ca = conv(a, ct), cb = conv(b, ct);
        if (ct == Tag::FLOAT32) { return mk(std::fmin(ca.f, cb.f)); }
        return binary_lane(BinaryOp::LESS, ca, cb).b ? ca : cb;
    }
    static Lane lane_max(const Lane &a, const Lane &b) {
        auto ct = promote(a.k, b.k);
        auto ca = conv(a, ct), cb = conv(b, ct);
        if (ct == Tag::FLOAT32) { return mk(std::fmax(ca.f, cb.f)); }
        return binary_lane(BinaryOp::GREATER, ca, cb).b ? ca : cb;
    }
    static float fdot(const std::vector<Lane> &a, const std::vector<Lane> &b) {// a.x*b.x + a.y*b.y + ... (:3483)
        auto acc = a[0].f * b[0].f;
        for (auto i = 1u; i < a.size(); i++) { acc = acc + a[i].f * b[i].f; }
        return acc;
    }

    Val make_vector(const CallExpr *e) {
        auto s = shape_of(e->type());
        std::vector<Lane> all;
        for (auto arg : e->arguments()) {
            auto ls = lanes(eval(arg));
            all.insert(all.end(), ls.begin(), ls.end());
        }
        std::vector<Lane> out;
        if (all.size() == 1u) {
            out.assign(s.n, all[0]);
        } else {
            if (all.size() < s.n) { fail("make_vector: too few components"); }
            out.assign(all.begin(), all.begin() + s.n);// truncation of a wider vector
        }
        return from_lanes(e->type(), out);
    }

    Val make_matrix(const CallExpr *e) {
        auto n = e->type()->dimension();
        auto args = e->arguments();
        std::vector<Lane> out(n * n, mk(0.f));
        if (args.size() == 1u && args[0]->type()->is_matrix()) {
            auto src = lanes(eval(args[0]));
            auto m = args[0]->type()->dimension();
            for (auto c = 0u; c < n; c++) {
                for (auto r = 0u; r < n; r++) { out[c * n + r] = (c < m && r < m) ? src[c * m + r] : mk(c == r ? 1.f : 0.f); }
            }
            return from_lanes(e->type(), out);
        }
        std::vector<Lane> all;
        for (auto arg : args) {
            auto ls = lanes(eval(arg));
            all.insert(all.end(), ls.begin(), ls.end());
        }
        if (all.size() != n * n) { fail("make_matrix: component count"); }
        return from_lanes(e->type(), all);
    }

    Val call_custom(const CallExpr *e) {
        auto callee = e->custom();
        Machine m{callee};
        auto params = callee.arguments();
        auto args = e->arguments();
        if (params.size() != args.size()) { fail("callable argument count mismatch"); }
        std::vector<Val> temporaries;
        temporaries.reserve(args.size());
        for (auto i = 0u; i < params.size(); i++) {
            auto p = params[i];
            if (p.is_resource()) {
                if (p.tag() != Variable::Tag::BUFFER) { fail("only buffer resources are supported"); }
                m.bind_buffer(p, buffer_of(args[i]));
            } else if (p.is_reference()) {
                Ptr lv;
                if (try_lvalue(args[i], lv)) {
                    m.bind_reference(p, lv.p);
                } else {// reference to a temporary
                    temporaries.push_back(eval(args[i]));
                    m.bind_reference(p, temporaries.back().p());
                }
            } else {
                auto v = eval(args[i]);
                m.bind_value(p, v.p());
            }
        }
        return m.run();
    }

    Val call(const CallExpr *e) {
        auto op = e->op();
        auto args = e->arguments();
        auto F = [&](size_t i) { return lanes(eval(args[i])); };
        switch (op) {
            case CallOp::CUSTOM: return call_custom(e);
            case CallOp::ALL: {
                auto a = F(0);
                auto r = true;
                for (auto &x : a) { r = r && x.b; }
                return from_lanes(e->type(), {mk(r)});
            }
            case CallOp::ANY: {
                auto a = F(0);
                auto r = false;
                for (auto &x : a) { r = r || x.b; }
                return from_lanes(e->type(), {mk(r)});
            }
            case CallOp::SELECT: {// select(f, t, p) = p ? t : f
                auto fv = eval(args[0]);
                auto tv = eval(args[1]);
                auto p = F(2);
                if (p.size() == 1u) { return p[0].b ? tv : fv; }
                auto lf = lanes(fv), lt = lanes(tv);
                std::vector<Lane> out;
                for (auto i = 0u; i < p.size(); i++) { out.push_back(p[i].b ? lt[lt.size() == 1u ? 0u : i] : lf[lf.size() == 1u ? 0u : i]); }
                return from_lanes(e->type(), out);
            }
            case CallOp::CLAMP: {// min(max(v, lo), hi)
                auto v = F(0), lo = F(1), hi = F(2);
                std::vector<Lane> out;
                for (auto i = 0u; i < v.size(); i++) {
                    out.push_back(lane_min(lane_max(v[i], lo[lo.size() == 1u ? 0u : i]), hi[hi.size() == 1u ? 0u : i]));
                }
                return from_lanes(e->type(), out);
            }
            case CallOp::SATURATE: return map_f(e, [](float x) { return std::fmin(std::fmax(x, 0.f), 1.f); });
            case CallOp::LERP: return map_fff(e, [](float a, float b, float t) { return t * (b - a) + a; });
            case CallOp::SMOOTHSTEP:
                return map_fff(e, [](float e0, float e1, float x) {
                    auto t = std::fmin(std::fmax((x - e0) / (e1 - e0), 0.f), 1.f);
                    return t * t * (3.f - 2.f * t);
                });
            case CallOp::STEP: return map_ff(e, [](float edge, float x) { return x < edge ? 0.f : 1.f; });
            case CallOp::ABS: {
                auto a = F(0);
                std::vector<Lane> out;
                for (auto &x : a) {
                    out.push_back(visit(x, [](auto v) -> Lane {
                        using T = decltype(v);
                        if constexpr (std::is_floating_point_v<T>) { return mk(static_cast<T>(std::fabs(v))); }
                        else if constexpr (std::is_signed_v<T>) { return mk(static_cast<T>(v < 0 ? -v : v)); }
                        else { return mk(v); }
                    }));
                }
                return from_lanes(e->type(), out);
            }
            case CallOp::MIN:
            case CallOp::MAX: {
                auto a = F(0), b = F(1);
                auto n = std::max(a.size(), b.size());
                std::vector<Lane> out;
                for (auto i = 0u; i < n; i++) {
                    auto &x = a[a.size() == 1u ? 0u : i];
                    auto &y = b[b.size() == 1u ? 0u : i];
                    out.push_back(op == CallOp::MIN ? lane_min(x, y) : lane_max(x, y));
                }
                return from_lanes(e->type(), out);
            }
            case CallOp::CLZ:
            case CallOp::CTZ:
            case CallOp::POPCOUNT:
            case CallOp::REVERSE: {
                auto a = F(0);
                std::vector<Lane> out;
                for (auto &x : a) {
                    auto u = conv(x, Tag::UINT32).u;
                    uint32_t r = 0u;
                    if (op == CallOp::CLZ) { r = u == 0u ? 32u : static_cast<uint32_t>(__builtin_clz(u)); }
                    else if (op == CallOp::CTZ) { r = u == 0u ? 32u : static_cast<uint32_t>(__builtin_ctz(u)); }
                    else if (op == CallOp::POPCOUNT) { r = static_cast<uint32_t>(__builtin_popcount(u)); }
                    else {
                        for (auto i = 0u; i < 32u; i++) { r |= ((u >> i) & 1u) << (31u - i); }
                    }
                    out.push_back(mk(r));
                }
                return from_lanes(e->type(), out);
            }
            case CallOp::ISINF:
            case CallOp::ISNAN: {
                auto a = F(0);
                std::vector<Lane> out;
                for (auto &x : a) { out.push_back(mk(op == CallOp::ISINF ? isinf_impl(x.f) : isnan_impl(x.f))); }
                return from_lanes(e->type(), out);
            }
            case CallOp::ACOS: return map_f(e, [](float x) { return std::acos(x); });
            case CallOp::ACOSH: return map_f(e, [](float x) { return std::acosh(x); });
            case CallOp::ASIN: return map_f(e, [](float x) { return std::asin(x); });
            case CallOp::ASINH: return map_f(e, [](float x) { return std::asinh(x); });
            case CallOp::ATAN: return map_f(e, [](float x) { return std::atan(x); });
            case CallOp::ATAN2: return map_ff(e, [](float y, float x) { return std::atan2(y, x); });
            case CallOp::ATANH: return map_f(e, [](float x) { return std::atanh(x); });
            case CallOp::COS: return map_f(e, [](float x) { return std::cos(x); });
            case CallOp::COSH: return map_f(e, [](float x) { return std::cosh(x); });
            case CallOp::SIN: return map_f(e, [](float x) { return std::sin(x); });
            case CallOp::SINH: return map_f(e, [](float x) { return std::sinh(x); });
            case CallOp::TAN: return map_f(e, [](float x) { return std::tan(x); });
            case CallOp::TANH: return map_f(e, [](float x) { return std::tanh(x); });
            case CallOp::EXP: return map_f(e, [](float x) { return std::exp(x); });
            case CallOp::EXP2: return map_f(e, [](float x) { return std::exp2(x); });
            case CallOp::EXP10: return map_f(e, [](float x) { return std::pow(10.f, x); });
            case CallOp::LOG: return map_f(e, [](float x) { return std::log(x); });
            case CallOp::LOG2: return map_f(e, [](float x) { return std::log2(x); });
            case CallOp::LOG10: return map_f(e, [](float x) { return std::log10(x); });
            case CallOp::POW: return map_ff(e, [](float x, float y) { return powf_impl(x, y); });
            case CallOp::SQRT: return map_f(e, [](float x) { return std::sqrt(x); });
            case CallOp::RSQRT: return map_f(e, [](float x) { return 1.0f / std::sqrt(x); });
            case CallOp::CEIL: return map_f(e, [](float x) { return std::ceil(x); });
            case CallOp::FLOOR: return map_f(e, [](float x) { return std::floor(x); });
            case CallOp::FRACT: return map_f(e, [](float x) { return x - std::floor(x); });
            case CallOp::TRUNC: return map_f(e, [](float x) { return std::trunc(x); });
            case CallOp::ROUND: return map_f(e, [](float x) { return std::round(x); });
            case CallOp::FMA: return map_fff(e, [](float a, float b, float c) { return std::fma(a, b, c); });
            case CallOp::COPYSIGN: return map_ff(e, [](float a, float b) { return std::copysign(a, b); });
            case CallOp::CROSS: {
                auto u = F(0), v = F(1);
                return from_lanes(e->type(), {mk(u[1].f * v[2].f - v[1].f * u[2].f),
                                              mk(u[2].f * v[0].f - v[2].f * u[0].f),
                                              mk(u[0].f * v[1].f - v[0].f * u[1].f)});
            }
            case CallOp::DOT: {
                auto a = F(0), b = F(1);
                return from_lanes(e->type(), {mk(fdot(a, b))});
            }
            case CallOp::LENGTH: {
                auto a = F(0);
                return from_lanes(e->type(), {mk(std::sqrt(fdot(a, a)))});
            }
            case CallOp::LENGTH_SQUARED: {
                auto a = F(0);
                return from_lanes(e->type(), {mk(fdot(a, a))});
            }
            case CallOp::NORMALIZE: {// v * rsqrt(dot(v, v))
                auto a = F(0);
                auto s = 1.0f / std::sqrt(fdot(a, a));
                std::vector<Lane> out;
                for (auto &x : a) { out.push_back(mk(x.f * s)); }
                return from_lanes(e->type(), out);
            }
            case CallOp::FACEFORWARD: {// select(-n, n, dot(n_ref, i) < 0)
                auto n = F(0), i = F(1), nref = F(2);
                auto keep = fdot(nref, i) < 0.f;
                std::vector<Lane> out;
                for (auto &x : n) { out.push_back(mk(keep ? x.f : -x.f)); }
                return from_lanes(e->type(), out);
            }
            case CallOp::REFLECT: {// v - 2 * dot(v, n) * n
                auto v = F(0), n = F(1);
                auto s = 2.0f * fdot(v, n);
                std::vector<Lane> out;
                for (auto i = 0u; i < v.size(); i++) { out.push_back(mk(v[i].f - s * n[i].f)); }
                return from_lanes(e->type(), out);
            }
            case CallOp::REDUCE_SUM:
            case CallOp::REDUCE_PRODUCT:
            case CallOp::REDUCE_MIN:
            case CallOp::REDUCE_MAX: {
                auto a = F(0);
                auto acc = a[0];
                for (auto i = 1u; i < a.size(); i++) {
                    if (op == CallOp::REDUCE_SUM) { acc = binary_lane(BinaryOp::ADD, acc, a[i]); }
                    else if (op == CallOp::REDUCE_PRODUCT) { acc = binary_lane(BinaryOp::MUL, acc, a[i]); }
                    else if (op == CallOp::REDUCE_MIN) { acc = lane_min(acc, a[i]); }
                    else { acc = lane_max(acc, a[i]); }
                }
                return from_lanes(e->type(), {acc});
            }
            case CallOp::TRANSPOSE: {
                auto a = F(0);
                auto n = e->type()->dimension();
                std::vector<Lane> out(n * n);
                for (auto c = 0u; c < n; c++) {
                    for (auto r = 0u; r < n; r++) { out[c * n + r] = a[r * n + c]; }
                }
                return from_lanes(e->type(), out);
            }
            case CallOp::BUFFER_READ: {
                auto idx = static_cast<size_t>(conv(F(1)[0], Tag::UINT32).u);
                return buffer_read(args[0], idx, e->type());
            }
            case CallOp::BUFFER_WRITE: {
                auto b = buffer_of(args[0]);
                auto idx = static_cast<size_t>(conv(F(1)[0], Tag::UINT32).u);
                auto v = eval(args[2]);
                if (idx >= b.count) { fail("buffer write out of range"); }
                std::memcpy(b.data + idx * align_up(v.t->size(), v.t->alignment()), v.p(), v.m.size());
                return Val{};
            }
            case CallOp::BUFFER_SIZE: return from_lanes(e->type(), {mk(static_cast<uint64_t>(buffer_of(args[0]).count))});
            case CallOp::MAKE_BOOL2:
            case CallOp::MAKE_BOOL3:
            case CallOp::MAKE_BOOL4:
            case CallOp::MAKE_INT2:
            case CallOp::MAKE_INT3:
            case CallOp::MAKE_INT4:
            case CallOp::MAKE_UINT2:
            case CallOp::MAKE_UINT3:
            case CallOp::MAKE_UINT4:
            case CallOp::MAKE_FLOAT2:
            case CallOp::MAKE_FLOAT3:
            case CallOp::MAKE_FLOAT4:
            case CallOp::MAKE_LONG2:
            case CallOp::MAKE_LONG3:
            case CallOp::MAKE_LONG4:
            case CallOp::MAKE_ULONG2:
            case CallOp::MAKE_ULONG3:
            case CallOp::MAKE_ULONG4: return make_vector(e);
            case CallOp::MAKE_FLOAT2X2:
            case CallOp::MAKE_FLOAT3X3:
            case CallOp::MAKE_FLOAT4X4: return make_matrix(e);
            case CallOp::ASSERT:
            case CallOp::ASSUME: return Val{};
            case CallOp::UNREACHABLE: fail("unreachable() executed");
            case CallOp::ZERO: return Val{e->type()};
            case CallOp::ONE: {
                auto s = shape_of(e->type());
                std::vector<Lane> out(s.matrix ? s.n * s.n : s.n, conv(mk(1), s.elem));
                return from_lanes(e->type(), out);
            }
            default: fail("unsupported builtin call op " + std::to_string(static_cast<uint32_t>(op)));
        }
    }

    /* ---- statements ---- */
    void assign(const Expression *lhs, const Val &v) {
        Ptr lv;
        if (try_lvalue(lhs, lv)) {
            if (lv.t->size() != v.m.size()) {// scalar -> vector broadcast etc. never happens in LC; convert by lanes
                auto c = from_lanes(lv.t, lanes(v));
                std::memcpy(lv.p, c.p(), c.m.size());
            } else if (lv.t != v.t && (lv.t->is_scalar() || lv.t->is_vector())) {
                auto c = from_lanes(lv.t, lanes(v));
                std::memcpy(lv.p, c.p(), c.m.size());
            } else {
                std::memcpy(lv.p, v.p(), v.m.size());
            }
            return;
        }
        if (lhs->tag() == Expression::Tag::MEMBER) {// multi-component swizzle store
            auto m = static_cast<const MemberExpr *>(lhs);
            Ptr self;
            if (m->is_swizzle() && try_lvalue(m->self(), self)) {
                auto es = scalar_size(self.t->element()->tag());
                for (auto i = 0u; i < m->swizzle_size(); i++) { std::memcpy(self.p + m->swizzle_index(i) * es, v.p() + i * es, es); }
                return;
            }
        }
        fail("assignment to a non-lvalue");
    }

    Flow exec(const ScopeStmt *scope) {
        for (auto s : scope->statements()) {
            auto flow = exec(s);
            if (flow != Flow::NORMAL) { return flow; }
        }
        return Flow::NORMAL;
    }

    bool truth(const Expression *e) { return conv(lanes(eval(e))[0], Tag::BOOL).b; }

    Flow exec(const Statement *s) {
        switch (s->tag()) {
            case Statement::Tag::BREAK: return Flow::BREAK;
            case Statement::Tag::CONTINUE: return Flow::CONTINUE;
            case Statement::Tag::RETURN: {
                auto r = static_cast<const ReturnStmt *>(s);
                if (r->expression() != nullptr) { _ret = eval(r->expression()); }
                return Flow::RETURN;
            }
            case Statement::Tag::SCOPE: return exec(static_cast<const ScopeStmt *>(s));
            case Statement::Tag::IF: {
                auto i = static_cast<const IfStmt *>(s);
                return truth(i->condition()) ? exec(i->true_branch()) : exec(i->false_branch());
            }
            case Statement::Tag::LOOP: {
                auto l = static_cast<const LoopStmt *>(s);
                for (auto iter = 0u;; iter++) {
                    if (iter > (1u << 24u)) { fail("loop does not terminate"); }
                    auto flow = exec(l->body());
                    if (flow == Flow::BREAK) { break; }
                    if (flow == Flow::RETURN) { return flow; }
                }
                return Flow::NORMAL;
            }
            case Statement::Tag::EXPR: {
                (void)eval(static_cast<const ExprStmt *>(s)->expression());
                return Flow::NORMAL;
            }
            case Statement::Tag::SWITCH: {
                auto sw = static_cast<const SwitchStmt *>(s);
                auto value = conv(lanes(eval(sw->expression()))[0], Tag::INT64).l;
                const ScopeStmt *chosen = nullptr;
                const ScopeStmt *fallback = nullptr;
                for (auto c : sw->body()->statements()) {
                    if (c->tag() == Statement::Tag::SWITCH_CASE) {
                        auto sc = static_cast<const SwitchCaseStmt *>(c);
                        if (chosen == nullptr && conv(lanes(eval(sc->expression()))[0], Tag::INT64).l == value) { chosen = sc->body(); }
                    } else if (c->tag() == Statement::Tag::SWITCH_DEFAULT) {
                        fallback = static_cast<const SwitchDefaultStmt *>(c)->body();
                    }
                }
                if (chosen == nullptr) { chosen = fallback; }
                if (chosen == nullptr) { return Flow::NORMAL; }
                auto flow = exec(chosen);
                return flow == Flow::BREAK ? Flow::NORMAL : flow;
            }
            case Statement::Tag::ASSIGN: {
                auto a = static_cast<const AssignStmt *>(s);
                assign(a->lhs(), eval(a->rhs()));
                return Flow::NORMAL;
            }
            case Statement::Tag::FOR: {// for (; cond; var += step) body
                auto f = static_cast<const ForStmt *>(s);
                for (auto iter = 0u;; iter++) {
                    if (iter > (1u << 24u)) { fail("for loop does not terminate"); }
                    if (!truth(f->condition())) { break; }
                    auto flow = exec(f->body());
                    if (flow == Flow::BREAK) { break; }
                    if (flow == Flow::RETURN) { return flow; }
                    auto next = binary(BinaryOp::ADD, eval(f->variable()), eval(f->step()), f->variable()->type());
                    assign(f->variable(), next);
                }
                return Flow::NORMAL;
            }
            case Statement::Tag::COMMENT: return Flow::NORMAL;
            case Statement::Tag::PRINT: return Flow::NORMAL;
            default: fail("unsupported statement tag " + std::to_string(static_cast<int>(s->tag())));
        }
    }
};

}// namespace

std::vector<std::byte> call(Function f, std::vector<Arg> &args) {
    Machine m{f};
    auto params = f.arguments();
    if (params.size() != args.size()) { fail("entry argument count mismatch"); }
    for (auto i = 0u; i < params.size(); i++) {
        auto p = params[i];
        if (p.is_resource()) {
            if (p.tag() != Variable::Tag::BUFFER) { fail("only buffer resources are supported"); }
            m.bind_buffer(p, args[i].buffer);
        } else {
            if (args[i].bytes.size() != p.type()->size()) { fail("entry argument size mismatch"); }
            if (p.is_reference()) { m.bind_reference(p, args[i].bytes.data()); }
            else { m.bind_value(p, args[i].bytes.data()); }
        }
    }
    auto r = m.run();
    return r.m;
}

}// namespace refinterp
