// interp.cpp — host interpreter for LuisaCompute ASTs; see interp.h.  TEST INFRASTRUCTURE ONLY.
#include "interp.h"

#include <cmath>
#include <cstring>
#include <stdexcept>
#include <string>
#include <deque>
#include <mutex>
#include <thread>
#include <atomic>
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

/* a typed value; up to 64 bytes (any scalar / vector / matrix, small structs) live inline */
struct Val {
    const Type *t{nullptr};
    size_t n{0u};
    alignas(16) std::byte inl[64];
    std::vector<std::byte> heap;
    Val() = default;
    explicit Val(const Type *type) : t{type}, n{type == nullptr ? 0u : type->size()} {
        if (n <= sizeof(inl)) { std::memset(inl, 0, sizeof(inl)); } else { heap.assign(n, std::byte{0}); }
    }
    std::byte *p() { return n <= sizeof(inl) ? inl : heap.data(); }
    const std::byte *p() const { return n <= sizeof(inl) ? inl : heap.data(); }
    size_t size() const { return n; }
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

/* up to 16 lanes (float4x4) without heap traffic */
struct Lanes {
    Lane v[16];
    uint32_t n{0u};
    Lanes() = default;
    Lanes(std::initializer_list<Lane> init) { for (auto &x : init) { v[n++] = x; } }
    Lanes(size_t count, const Lane &x) { for (n = 0u; n < count; n++) { v[n] = x; } }
    void push_back(const Lane &x) { if (n >= 16u) { fail("more than 16 lanes"); } v[n++] = x; }
    [[nodiscard]] size_t size() const { return n; }
    Lane &operator[](size_t i) { return v[i]; }
    const Lane &operator[](size_t i) const { return v[i]; }
    Lane *begin() { return v; }
    Lane *end() { return v + n; }
    const Lane *begin() const { return v; }
    const Lane *end() const { return v + n; }
};

Lanes lanes(const Val &v) {
    auto s = shape_of(v.t);
    Lanes out;
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

Val from_lanes(const Type *t, const Lanes &ls) {
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
        Lanes out;
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
    Lanes out;
    for (auto i = 0u; i < n; i++) { out.push_back(binary_lane(op, la[sa.n == 1u ? 0u : i], lb[sb.n == 1u ? 0u : i])); }
    (void)is_relational;
    return from_lanes(rt, out);
}

Val unary(UnaryOp op, const Val &a, const Type *rt) {
    auto la = lanes(a);
    Lanes out;
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
    enum struct Kind : uint8_t { UNSET, DATA, BUFFER, BINDLESS_ARRAY, ACCEL } kind{Kind::UNSET};
    std::byte *ptr{nullptr};
    alignas(16) std::byte inl[64];
    std::vector<std::byte> own;
    BufferArg buffer;
    uint64_t handle{0u};
};

/* atomics on buffers shared by the host threads of one launch: striped locks are plenty here */
std::mutex &atomic_lock(const void *address) {
    static std::mutex locks[256];
    return locks[(reinterpret_cast<uintptr_t>(address) >> 2u) & 255u];
}

enum struct Flow { NORMAL, BREAK, CONTINUE, RETURN };

class Machine {

private:
    Function _f;
    DeviceResources *_res{nullptr};
    std::deque<Slot> _slots;// indexed by variable uid (uids are small and dense per function)
    Val _ret;
    /* ray queries (RayQueryAll / RayQueryAny variables of this function), keyed by variable uid */
    struct QueryState {
        uint64_t accel{0u};
        RayData ray{};
        uint32_t mask{0xffu};
        bool any{false};
        bool terminated{false};
        HitData committed{~0u, ~0u, {0.f, 0.f}, 0.f, 0u};
        HitData candidate{~0u, ~0u, {0.f, 0.f}, 0.f, 0u};
    };
    std::unordered_map<uint32_t, QueryState> _queries;
    QueryState &query_of(const Expression *e) {
        if (e->tag() != Expression::Tag::REF) { fail("ray query expression is not a variable"); }
        auto it = _queries.find(static_cast<const RefExpr *>(e)->variable().uid());
        if (it == _queries.end()) { fail("ray query used before it was created"); }
        return it->second;
    }

    Slot &raw_slot(uint32_t uid) {
        if (uid >= _slots.size()) { _slots.resize(uid + 1u); }
        return _slots[uid];
    }

public:
    Machine(Function f, DeviceResources *res) : _f{f}, _res{res} {}
    Slot &slot(Variable v) {
        auto &s = raw_slot(v.uid());
        if (s.kind == Slot::Kind::UNSET) {
            if (v.is_resource()) { fail("unbound resource variable"); }
            s.kind = Slot::Kind::DATA;
            auto n = std::max<size_t>(v.type()->size(), 1u);// locals are zero-initialised
            if (n <= sizeof(s.inl)) {
                std::memset(s.inl, 0, sizeof(s.inl));
                s.ptr = s.inl;
            } else {
                s.own.assign(n, std::byte{0});
                s.ptr = s.own.data();
            }
        }
        return s;
    }
    void bind_value(Variable v, const std::byte *data) {
        auto &s = slot(v);
        std::memcpy(s.ptr, data, v.type()->size());
    }
    void bind_reference(Variable v, std::byte *target) {
        auto &s = raw_slot(v.uid());
        s.kind = Slot::Kind::DATA;
        s.ptr = target;
    }
    void bind_buffer(Variable v, BufferArg b) {
        auto &s = raw_slot(v.uid());
        s.kind = Slot::Kind::BUFFER;
        s.buffer = b;
    }
    void bind_handle(Variable v, Slot::Kind kind, uint64_t handle) {
        auto &s = raw_slot(v.uid());
        s.kind = kind;
        s.handle = handle;
    }
    void bind_arg(Variable p, const Arg &a) {
        switch (a.kind) {
            case Arg::Kind::BUFFER: bind_buffer(p, a.buffer); break;
            case Arg::Kind::BINDLESS_ARRAY: bind_handle(p, Slot::Kind::BINDLESS_ARRAY, a.handle); break;
            case Arg::Kind::ACCEL: bind_handle(p, Slot::Kind::ACCEL, a.handle); break;
            default:
                if (a.bytes.size() != p.type()->size()) { fail("argument size mismatch"); }
                bind_value(p, a.bytes.data());
                break;
        }
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
                if (v.tag() != Variable::Tag::LOCAL && v.tag() != Variable::Tag::REFERENCE && v.tag() != Variable::Tag::SHARED &&
                    raw_slot(v.uid()).kind != Slot::Kind::DATA) {
                    fail("builtin variable (thread / dispatch id) used outside a kernel launch");
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
            std::memcpy(v.p(), lv.p, v.size());
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
                    std::memcpy(out.p(), self.p() + member_offset(self.t, m->member_index()), out.size());
                }
                return out;
            }
            case Expression::Tag::ACCESS: {
                auto a = static_cast<const AccessExpr *>(e);
                auto idx = static_cast<size_t>(conv(lanes(eval(a->index()))[0], Tag::UINT32).u);
                if (a->range()->type()->is_buffer()) { return buffer_read(a->range(), idx, e->type()); }
                auto range = eval(a->range());
                Val out{e->type()};
                std::memcpy(out.p(), range.p() + element_offset(range.t, idx), out.size());
                return out;
            }
            case Expression::Tag::LITERAL: {
                auto l = static_cast<const LiteralExpr *>(e);
                Val out{e->type()};
                luisa::visit(
                    [&out](auto v) {
                        if (sizeof(v) != out.size()) { fail("literal size mismatch"); }
                        std::memcpy(out.p(), &v, sizeof(v));
                    },
                    l->value().to_variant());
                return out;
            }
            case Expression::Tag::CONSTANT: {
                auto c = static_cast<const ConstantExpr *>(e);
                Val out{e->type()};
                std::memcpy(out.p(), c->data().raw(), out.size());
                return out;
            }
            case Expression::Tag::CAST: {
                auto c = static_cast<const CastExpr *>(e);
                auto src = eval(c->expression());
                if (c->op() == CastOp::BITWISE) {
                    Val out{e->type()};
                    if (out.size() != src.size()) { fail("bitwise cast between different sizes"); }
                    std::memcpy(out.p(), src.p(), out.size());
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
        auto &s = raw_slot(v.uid());
        if (s.kind != Slot::Kind::BUFFER || s.buffer.data == nullptr) { fail("unbound buffer argument"); }
        return s.buffer;
    }
    uint64_t handle_of(const Expression *e, Slot::Kind kind) {
        if (e->tag() != Expression::Tag::REF) { fail("resource expression is not a variable"); }
        auto &s = raw_slot(static_cast<const RefExpr *>(e)->variable().uid());
        if (s.kind != kind) { fail("resource variable bound to the wrong kind"); }
        return s.handle;
    }
    DeviceResources &res() {
        if (_res == nullptr) { fail("device resources used outside a device launch"); }
        return *_res;
    }

    Val buffer_read(const Expression *buffer, size_t idx, const Type *elem) {
        return read_element(buffer_of(buffer), idx, elem);
    }
    Val read_element(BufferArg b, size_t idx, const Type *elem) {
        auto stride = align_up(elem->size(), elem->alignment());
        if ((idx + 1u) * stride > b.size_bytes) { fail("buffer read out of range"); }
        Val out{elem};
        std::memcpy(out.p(), b.data + idx * stride, out.size());
        return out;
    }

    /* ---- calls ---- */
    template<typename F>
    Val map_f(const CallExpr *e, F &&fn) {// float -> float, per lane
        auto a = lanes(eval(e->arguments()[0]));
        Lanes out;
        for (auto &x : a) { out.push_back(mk(static_cast<float>(fn(conv(x, Tag::FLOAT32).f)))); }
        return from_lanes(e->type(), out);
    }
    template<typename F>
    Val map_ff(const CallExpr *e, F &&fn) {
        auto a = lanes(eval(e->arguments()[0]));
        auto b = lanes(eval(e->arguments()[1]));
        auto n = std::max(a.size(), b.size());
        Lanes out;
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
        Lanes out;
        auto at = [](const Lanes &v, size_t i) { return conv(v[v.size() == 1u ? 0u : i], Tag::FLOAT32).f; };
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
    static float fdot(const Lanes &a, const Lanes &b) {// a.x*b.x + a.y*b.y + ... (:3483)
        auto acc = a[0].f * b[0].f;
        for (auto i = 1u; i < a.size(); i++) { acc = acc + a[i].f * b[i].f; }
        return acc;
    }

    /* atomic ops: arguments = (buffer, index, [member / component indices ...], value[s]) */
    Val atomic(const CallExpr *e) {
        auto args = e->arguments();
        auto n_values = e->op() == CallOp::ATOMIC_COMPARE_EXCHANGE ? 2u : 1u;
        if (!args[0]->type()->is_buffer()) { fail("atomics are implemented for buffers only"); }
        auto b = buffer_of(args[0]);
        auto t = args[0]->type()->element();
        auto idx = static_cast<size_t>(conv(lanes(eval(args[1]))[0], Tag::UINT32).u);
        auto stride = align_up(t->size(), t->alignment());
        if ((idx + 1u) * stride > b.size_bytes) { fail("atomic access out of range"); }
        auto p = b.data + idx * stride;
        for (auto i = 2u; i + n_values < args.size(); i++) {
            auto k = static_cast<size_t>(conv(lanes(eval(args[i]))[0], Tag::UINT32).u);
            if (t->is_structure()) {
                p += member_offset(t, static_cast<uint32_t>(k));
                t = t->members()[k];
            } else {
                p += element_offset(t, k);
                t = t->is_matrix() ? Type::vector(Type::of<float>(), t->dimension()) : t->element();
            }
        }
        if (!t->is_scalar()) { fail("atomic on a non-scalar"); }
        auto v0 = conv(lanes(eval(args[args.size() - n_values]))[0], t->tag());
        std::scoped_lock lock{atomic_lock(p)};
        auto old = load(p, t->tag());
        Lane next = old;
        switch (e->op()) {
            case CallOp::ATOMIC_EXCHANGE: next = v0; break;
            case CallOp::ATOMIC_COMPARE_EXCHANGE: {
                auto desired = conv(lanes(eval(args[args.size() - 1u]))[0], t->tag());
                if (binary_lane(BinaryOp::EQUAL, old, v0).b) { next = desired; }
                break;
            }
            case CallOp::ATOMIC_FETCH_ADD: next = binary_lane(BinaryOp::ADD, old, v0); break;
            case CallOp::ATOMIC_FETCH_SUB: next = binary_lane(BinaryOp::SUB, old, v0); break;
            case CallOp::ATOMIC_FETCH_AND: next = binary_lane(BinaryOp::BIT_AND, old, v0); break;
            case CallOp::ATOMIC_FETCH_OR: next = binary_lane(BinaryOp::BIT_OR, old, v0); break;
            case CallOp::ATOMIC_FETCH_XOR: next = binary_lane(BinaryOp::BIT_XOR, old, v0); break;
            case CallOp::ATOMIC_FETCH_MIN: next = lane_min(old, v0); break;
            default: next = lane_max(old, v0); break;
        }
        store(p, conv(next, t->tag()));
        return from_lanes(e->type(), {old});
    }

    Val make_vector(const CallExpr *e) {
        auto s = shape_of(e->type());
        Lanes all;
        for (auto arg : e->arguments()) {
            for (auto &x : lanes(eval(arg))) { all.push_back(x); }
        }
        Lanes out;
        if (all.size() == 1u) {
            out = Lanes(s.n, all[0]);
        } else {
            if (all.size() < s.n) { fail("make_vector: too few components"); }
            for (auto i = 0u; i < s.n; i++) { out.push_back(all[i]); }// truncation of a wider vector
        }
        return from_lanes(e->type(), out);
    }

    Val make_matrix(const CallExpr *e) {
        auto n = e->type()->dimension();
        auto args = e->arguments();
        Lanes out(n * n, mk(0.f));
        if (args.size() == 1u && args[0]->type()->is_matrix()) {
            auto src = lanes(eval(args[0]));
            auto m = args[0]->type()->dimension();
            for (auto c = 0u; c < n; c++) {
                for (auto r = 0u; r < n; r++) { out[c * n + r] = (c < m && r < m) ? src[c * m + r] : mk(c == r ? 1.f : 0.f); }
            }
            return from_lanes(e->type(), out);
        }
        Lanes all;
        for (auto arg : args) {
            for (auto &x : lanes(eval(arg))) { all.push_back(x); }
        }
        if (all.size() != n * n) { fail("make_matrix: component count"); }
        return from_lanes(e->type(), all);
    }

    Val call_custom(const CallExpr *e) {
        auto callee = e->custom();
        Machine m{callee, _res};
        auto params = callee.arguments();
        auto args = e->arguments();
        if (params.size() != args.size()) { fail("callable argument count mismatch"); }
        std::vector<Val> temporaries;
        temporaries.reserve(args.size());
        for (auto i = 0u; i < params.size(); i++) {
            auto p = params[i];
            if (p.is_resource()) {
                if (args[i]->tag() != Expression::Tag::REF) { fail("resource argument is not a variable"); }
                auto &src = raw_slot(static_cast<const RefExpr *>(args[i])->variable().uid());
                if (src.kind == Slot::Kind::BUFFER) { m.bind_buffer(p, src.buffer); }
                else if (src.kind == Slot::Kind::BINDLESS_ARRAY || src.kind == Slot::Kind::ACCEL) { m.bind_handle(p, src.kind, src.handle); }
                else { fail("unsupported resource argument"); }
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
                Lanes out;
                for (auto i = 0u; i < p.size(); i++) { out.push_back(p[i].b ? lt[lt.size() == 1u ? 0u : i] : lf[lf.size() == 1u ? 0u : i]); }
                return from_lanes(e->type(), out);
            }
            case CallOp::CLAMP: {// min(max(v, lo), hi)
                auto v = F(0), lo = F(1), hi = F(2);
                Lanes out;
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
                Lanes out;
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
                Lanes out;
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
                Lanes out;
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
                Lanes out;
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
                Lanes out;
                for (auto &x : a) { out.push_back(mk(x.f * s)); }
                return from_lanes(e->type(), out);
            }
            case CallOp::FACEFORWARD: {// select(-n, n, dot(n_ref, i) < 0)
                auto n = F(0), i = F(1), nref = F(2);
                auto keep = fdot(nref, i) < 0.f;
                Lanes out;
                for (auto &x : n) { out.push_back(mk(keep ? x.f : -x.f)); }
                return from_lanes(e->type(), out);
            }
            case CallOp::REFLECT: {// v - 2 * dot(v, n) * n
                auto v = F(0), n = F(1);
                auto s = 2.0f * fdot(v, n);
                Lanes out;
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
            case CallOp::DETERMINANT:
            case CallOp::INVERSE: {// cofactor expansion as the CUDA backend's lc_determinant / lc_inverse (GLM's), cuda_device_math.h:3554-3680
                auto a = F(0);
                auto n = args[0]->type()->dimension();
                auto M = [&](uint32_t c, uint32_t r) { return a[c * n + r].f; };
                if (n == 2u) {
                    auto det = M(0, 0) * M(1, 1) - M(1, 0) * M(0, 1);
                    if (op == CallOp::DETERMINANT) { return from_lanes(e->type(), {mk(det)}); }
                    auto inv = 1.0f / det;
                    return from_lanes(e->type(), {mk(M(1, 1) * inv), mk(-M(0, 1) * inv), mk(-M(1, 0) * inv), mk(M(0, 0) * inv)});
                }
                if (n == 3u) {
                    auto det = M(0, 0) * (M(1, 1) * M(2, 2) - M(2, 1) * M(1, 2)) - M(1, 0) * (M(0, 1) * M(2, 2) - M(2, 1) * M(0, 2)) +
                               M(2, 0) * (M(0, 1) * M(1, 2) - M(1, 1) * M(0, 2));
                    if (op == CallOp::DETERMINANT) { return from_lanes(e->type(), {mk(det)}); }
                    auto inv = 1.0f / det;
                    return from_lanes(e->type(), {mk((M(1, 1) * M(2, 2) - M(2, 1) * M(1, 2)) * inv), mk((M(2, 1) * M(0, 2) - M(0, 1) * M(2, 2)) * inv),
                                                  mk((M(0, 1) * M(1, 2) - M(1, 1) * M(0, 2)) * inv), mk((M(2, 0) * M(1, 2) - M(1, 0) * M(2, 2)) * inv),
                                                  mk((M(0, 0) * M(2, 2) - M(2, 0) * M(0, 2)) * inv), mk((M(1, 0) * M(0, 2) - M(0, 0) * M(1, 2)) * inv),
                                                  mk((M(1, 0) * M(2, 1) - M(2, 0) * M(1, 1)) * inv), mk((M(2, 0) * M(0, 1) - M(0, 0) * M(2, 1)) * inv),
                                                  mk((M(0, 0) * M(1, 1) - M(1, 0) * M(0, 1)) * inv)});
                }
                // 4x4: 2x2 sub-determinants of rows (2,3), (1,3), (1,2) over column pairs, then the adjugate by columns
                auto sub = [&](uint32_t r0, uint32_t r1, uint32_t c0, uint32_t c1) { return M(c0, r0) * M(c1, r1) - M(c1, r0) * M(c0, r1); };
                float c00 = sub(2, 3, 2, 3), c02 = sub(2, 3, 1, 3), c03 = sub(2, 3, 1, 2);
                float c04 = sub(1, 3, 2, 3), c06 = sub(1, 3, 1, 3), c07 = sub(1, 3, 1, 2);
                float c08 = sub(1, 2, 2, 3), c10 = sub(1, 2, 1, 3), c11 = sub(1, 2, 1, 2);
                float c12 = sub(0, 3, 2, 3), c14 = sub(0, 3, 1, 3), c15 = sub(0, 3, 1, 2);
                float c16 = sub(0, 2, 2, 3), c18 = sub(0, 2, 1, 3), c19 = sub(0, 2, 1, 2);
                float c20 = sub(0, 1, 2, 3), c22 = sub(0, 1, 1, 3), c23 = sub(0, 1, 1, 2);
                float fac0[4] = {c00, c00, c02, c03}, fac1[4] = {c04, c04, c06, c07}, fac2[4] = {c08, c08, c10, c11};
                float fac3[4] = {c12, c12, c14, c15}, fac4[4] = {c16, c16, c18, c19}, fac5[4] = {c20, c20, c22, c23};
                float v0[4] = {M(1, 0), M(0, 0), M(0, 0), M(0, 0)}, v1[4] = {M(1, 1), M(0, 1), M(0, 1), M(0, 1)};
                float v2[4] = {M(1, 2), M(0, 2), M(0, 2), M(0, 2)}, v3_[4] = {M(1, 3), M(0, 3), M(0, 3), M(0, 3)};
                float inv_c[4][4];
                for (auto i = 0u; i < 4u; i++) {
                    auto sa = (i % 2u == 0u) ? 1.0f : -1.0f;
                    inv_c[0][i] = (v1[i] * fac0[i] - v2[i] * fac1[i] + v3_[i] * fac2[i]) * sa;
                    inv_c[1][i] = (v0[i] * fac0[i] - v2[i] * fac3[i] + v3_[i] * fac4[i]) * -sa;
                    inv_c[2][i] = (v0[i] * fac1[i] - v1[i] * fac3[i] + v3_[i] * fac5[i]) * sa;
                    inv_c[3][i] = (v0[i] * fac2[i] - v1[i] * fac4[i] + v2[i] * fac5[i]) * -sa;
                }
                auto det = M(0, 0) * inv_c[0][0] + M(0, 1) * inv_c[1][0] + M(0, 2) * inv_c[2][0] + M(0, 3) * inv_c[3][0];
                if (op == CallOp::DETERMINANT) { return from_lanes(e->type(), {mk(det)}); }
                auto inv = 1.0f / det;
                Lanes out;
                for (auto c = 0u; c < 4u; c++) {
                    for (auto r = 0u; r < 4u; r++) { out.push_back(mk(inv_c[c][r] * inv)); }
                }
                return from_lanes(e->type(), out);
            }
            case CallOp::TRANSPOSE: {
                auto a = F(0);
                auto n = e->type()->dimension();
                Lanes out(n * n, mk(0.f));
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
                auto stride = align_up(v.t->size(), v.t->alignment());
                if ((idx + 1u) * stride > b.size_bytes) { fail("buffer write out of range"); }
                std::memcpy(b.data + idx * stride, v.p(), v.size());
                return Val{};
            }
            case CallOp::BUFFER_SIZE: {
                auto elem = args[0]->type()->element();
                return from_lanes(e->type(), {mk(static_cast<uint64_t>(buffer_of(args[0]).size_bytes / align_up(elem->size(), elem->alignment())))});
            }
            case CallOp::BINDLESS_BUFFER_READ: {
                auto array = handle_of(args[0], Slot::Kind::BINDLESS_ARRAY);
                auto slot_index = conv(F(1)[0], Tag::UINT32).u;
                auto idx = static_cast<size_t>(conv(F(2)[0], Tag::UINT32).u);
                auto b = res().bindless_buffer(array, slot_index);
                if (b.data == nullptr) { return Val{e->type()}; }// unpopulated slot: see refdevice.cpp bindless_buffer
                return read_element(b, idx, e->type());
            }
            case CallOp::BINDLESS_TEXTURE2D_SAMPLE:
            case CallOp::BINDLESS_TEXTURE2D_SAMPLE_LEVEL: {// level 0 only (the image plugin samples without LOD, image.cpp:166)
                auto array = handle_of(args[0], Slot::Kind::BINDLESS_ARRAY);
                auto slot_index = conv(F(1)[0], Tag::UINT32).u;
                auto uv = F(2);
                float out[4];
                res().bindless_tex2d_sample(array, slot_index, uv[0].f, uv[1].f, out);
                return from_lanes(e->type(), {mk(out[0]), mk(out[1]), mk(out[2]), mk(out[3])});
            }
            case CallOp::BINDLESS_TEXTURE2D_READ: {
                auto array = handle_of(args[0], Slot::Kind::BINDLESS_ARRAY);
                auto slot_index = conv(F(1)[0], Tag::UINT32).u;
                auto xy = F(2);
                float out[4];
                res().bindless_tex2d_read(array, slot_index, conv(xy[0], Tag::UINT32).u, conv(xy[1], Tag::UINT32).u, out);
                return from_lanes(e->type(), {mk(out[0]), mk(out[1]), mk(out[2]), mk(out[3])});
            }
            case CallOp::BINDLESS_TEXTURE3D_READ: {// point reads of a volume: the PMJ02BN sampler's blue-noise textures
                auto array = handle_of(args[0], Slot::Kind::BINDLESS_ARRAY);
                auto slot_index = conv(F(1)[0], Tag::UINT32).u;
                auto xyz = F(2);
                float out[4];
                res().bindless_tex3d_read(array, slot_index, conv(xyz[0], Tag::UINT32).u, conv(xyz[1], Tag::UINT32).u, conv(xyz[2], Tag::UINT32).u, out);
                return from_lanes(e->type(), {mk(out[0]), mk(out[1]), mk(out[2]), mk(out[3])});
            }
            case CallOp::BINDLESS_TEXTURE2D_SIZE: {
                auto array = handle_of(args[0], Slot::Kind::BINDLESS_ARRAY);
                uint32_t size[2];
                res().bindless_tex2d_size(array, conv(F(1)[0], Tag::UINT32).u, size);
                return from_lanes(e->type(), {mk(size[0]), mk(size[1])});
            }
            case CallOp::RAY_TRACING_TRACE_CLOSEST:
            case CallOp::RAY_TRACING_TRACE_ANY: {
                auto accel = handle_of(args[0], Slot::Kind::ACCEL);
                auto ray = eval(args[1]);
                if (ray.size() != sizeof(RayData)) { fail("unexpected Ray layout"); }
                RayData rd;
                std::memcpy(&rd, ray.p(), sizeof(rd));
                auto mask = conv(F(2)[0], Tag::UINT32).u;
                if (op == CallOp::RAY_TRACING_TRACE_ANY) { return from_lanes(e->type(), {mk(res().trace_any(accel, rd, mask))}); }
                auto hit = res().trace_closest(accel, rd, mask);
                Val out{e->type()};
                if (out.size() > sizeof(HitData)) { fail("unexpected Hit layout"); }
                std::memcpy(out.p(), &hit, out.size());
                return out;
            }
            case CallOp::RAY_QUERY_WORLD_SPACE_RAY: {
                auto &q = query_of(args[0]);
                Val out{e->type()};
                if (out.size() != sizeof(RayData)) { fail("unexpected Ray layout"); }
                std::memcpy(out.p(), &q.ray, sizeof(RayData));
                return out;
            }
            case CallOp::RAY_QUERY_TRIANGLE_CANDIDATE_HIT: {
                auto &q = query_of(args[0]);
                Val out{e->type()};
                std::memcpy(out.p(), &q.candidate, std::min(out.size(), sizeof(HitData)));
                return out;
            }
            case CallOp::RAY_QUERY_COMMITTED_HIT: {// CommittedHit {inst, prim, bary, hit_type, committed_ray_t} (rtx/hit.h:18-24)
                auto &q = query_of(args[0]);
                struct { uint32_t inst, prim; float bary[2]; uint32_t hit_type; float t; } c{
                    q.committed.inst, q.committed.prim, {q.committed.bary[0], q.committed.bary[1]},
                    q.committed.inst == ~0u ? 0u : 1u, q.committed.committed_ray_t};
                Val out{e->type()};
                if (out.size() != sizeof(c)) { fail("unexpected CommittedHit layout"); }
                std::memcpy(out.p(), &c, sizeof(c));
                return out;
            }
            case CallOp::RAY_QUERY_COMMIT_TRIANGLE: {
                auto &q = query_of(args[0]);
                q.committed = q.candidate;
                q.ray.t_max = q.candidate.committed_ray_t;
                if (q.any) { q.terminated = true; }
                return Val{};
            }
            case CallOp::RAY_QUERY_TERMINATE: {
                query_of(args[0]).terminated = true;
                return Val{};
            }
            case CallOp::RAY_TRACING_INSTANCE_TRANSFORM: {
                auto accel = handle_of(args[0], Slot::Kind::ACCEL);
                float m[16];
                res().instance_transform(accel, conv(F(1)[0], Tag::UINT32).u, m);
                Lanes out;
                for (auto x : m) { out.push_back(mk(x)); }
                return from_lanes(e->type(), out);
            }
            case CallOp::ATOMIC_EXCHANGE:
            case CallOp::ATOMIC_COMPARE_EXCHANGE:
            case CallOp::ATOMIC_FETCH_ADD:
            case CallOp::ATOMIC_FETCH_SUB:
            case CallOp::ATOMIC_FETCH_AND:
            case CallOp::ATOMIC_FETCH_OR:
            case CallOp::ATOMIC_FETCH_XOR:
            case CallOp::ATOMIC_FETCH_MIN:
            case CallOp::ATOMIC_FETCH_MAX: return atomic(e);
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
                Lanes out(s.matrix ? s.n * s.n : s.n, conv(mk(1), s.elem));
                return from_lanes(e->type(), out);
            }
            default: fail("unsupported builtin call op " + std::to_string(static_cast<uint32_t>(op)));
        }
    }

    /* ---- statements ---- */
    void assign(const Expression *lhs, const Val &v) {
        Ptr lv;
        if (try_lvalue(lhs, lv)) {
            if (lv.t->size() != v.size()) {// scalar -> vector broadcast etc. never happens in LC; convert by lanes
                auto c = from_lanes(lv.t, lanes(v));
                std::memcpy(lv.p, c.p(), c.size());
            } else if (lv.t != v.t && (lv.t->is_scalar() || lv.t->is_vector())) {
                auto c = from_lanes(lv.t, lanes(v));
                std::memcpy(lv.p, c.p(), c.size());
            } else {
                std::memcpy(lv.p, v.p(), v.size());
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
                if (a->rhs()->tag() == Expression::Tag::CALL) {
                    auto c = static_cast<const CallExpr *>(a->rhs());
                    if (c->op() == CallOp::RAY_TRACING_QUERY_ALL || c->op() == CallOp::RAY_TRACING_QUERY_ANY) {
                        if (a->lhs()->tag() != Expression::Tag::REF) { fail("ray query assigned to a non-variable"); }
                        QueryState q;
                        q.accel = handle_of(c->arguments()[0], Slot::Kind::ACCEL);
                        auto ray = eval(c->arguments()[1]);
                        if (ray.size() != sizeof(RayData)) { fail("unexpected Ray layout"); }
                        std::memcpy(&q.ray, ray.p(), sizeof(RayData));
                        q.mask = conv(lanes(eval(c->arguments()[2]))[0], Tag::UINT32).u;
                        q.any = c->op() == CallOp::RAY_TRACING_QUERY_ANY;
                        q.committed.committed_ray_t = q.ray.t_max;
                        _queries[static_cast<const RefExpr *>(a->lhs())->variable().uid()] = q;
                        return Flow::NORMAL;
                    }
                }
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
            case Statement::Tag::RAY_QUERY: {
                auto rq = static_cast<const RayQueryStmt *>(s);
                auto &q = query_of(rq->query());
                for (auto &&c : res().candidates(q.accel, q.ray, q.mask)) {
                    if (q.terminated) { break; }
                    if (!(c.hit.committed_ray_t < q.ray.t_max)) { continue; }// a closer hit has been committed meanwhile
                    q.candidate = c.hit;
                    if (c.opaque) {// opaque geometry commits without the callback
                        q.committed = c.hit;
                        q.ray.t_max = c.hit.committed_ray_t;
                        if (q.any) { q.terminated = true; }
                    } else {
                        auto flow = exec(rq->on_triangle_candidate());
                        if (flow == Flow::RETURN) { return flow; }
                    }
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

std::vector<std::byte> call(Function f, std::vector<Arg> &args, DeviceResources *resources) {
    Machine m{f, resources};
    auto params = f.arguments();
    if (params.size() != args.size()) { fail("entry argument count mismatch"); }
    for (auto i = 0u; i < params.size(); i++) {
        auto p = params[i];
        if (!p.is_resource() && p.is_reference()) {
            if (args[i].bytes.size() != p.type()->size()) { fail("entry argument size mismatch"); }
            m.bind_reference(p, args[i].bytes.data());
        } else {
            m.bind_arg(p, args[i]);
        }
    }
    auto r = m.run();
    return std::vector<std::byte>(r.p(), r.p() + r.size());
}

void launch(Function f, const std::vector<Arg> &args, const uint32_t size[3], DeviceResources *resources, unsigned threads) {
    auto params = f.arguments();
    if (params.size() != args.size()) { fail("kernel argument count mismatch"); }
    auto total = static_cast<uint64_t>(size[0]) * size[1] * size[2];
    auto block = f.block_size();
    std::atomic<uint64_t> cursor{0u};
    std::mutex error_mutex;
    std::string error;
    auto worker = [&] {
        try {
            for (;;) {
                auto begin = cursor.fetch_add(64u);
                if (begin >= total) { break; }
                for (auto id = begin; id < std::min<uint64_t>(begin + 64u, total); id++) {
                    uint32_t d[4] = {static_cast<uint32_t>(id % size[0]), static_cast<uint32_t>(id / size[0] % size[1]),
                                     static_cast<uint32_t>(id / size[0] / size[1]), 0u};
                    uint32_t ds[4] = {size[0], size[1], size[2], 0u};
                    uint32_t tid[4] = {d[0] % block.x, d[1] % block.y, d[2] % block.z, 0u};
                    uint32_t bid[4] = {d[0] / block.x, d[1] / block.y, d[2] / block.z, 0u};
                    uint32_t zero[4] = {0u, 0u, 0u, 0u};
                    Machine m{f, resources};
                    for (auto i = 0u; i < params.size(); i++) { m.bind_arg(params[i], args[i]); }
                    for (auto v : f.builtin_variables()) {
                        switch (v.tag()) {
                            case Variable::Tag::DISPATCH_ID: m.bind_value(v, reinterpret_cast<const std::byte *>(d)); break;
                            case Variable::Tag::DISPATCH_SIZE: m.bind_value(v, reinterpret_cast<const std::byte *>(ds)); break;
                            case Variable::Tag::THREAD_ID: m.bind_value(v, reinterpret_cast<const std::byte *>(tid)); break;
                            case Variable::Tag::BLOCK_ID: m.bind_value(v, reinterpret_cast<const std::byte *>(bid)); break;
                            default: m.bind_value(v, reinterpret_cast<const std::byte *>(zero)); break;
                        }
                    }
                    (void)m.run();
                }
            }
        } catch (const std::exception &e) {
            std::scoped_lock lock{error_mutex};
            if (error.empty()) { error = e.what(); }
            cursor.store(total);
        }
    };
    // Small dispatches (the per-pass kernels of a small render, incl. the film's float atomics) run on ONE thread so that
    // the order of atomic accumulation - and with it the last bit of the film - is reproducible; large dispatches (e.g. the
    // 2048x1024 importance-map kernels of the Spherical environment) use every thread.
    if (total <= 32768u) { threads = 1u; }
    std::vector<std::thread> pool;
    for (auto i = 1u; i < std::max(threads, 1u); i++) { pool.emplace_back(worker); }
    worker();
    for (auto &t : pool) { t.join(); }
    if (!error.empty()) { throw std::runtime_error(error); }
}

}// namespace refinterp
