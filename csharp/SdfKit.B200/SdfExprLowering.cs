// SdfExprLowering.cs -- lowers an Expression<SdfFunc> (what SdfExprs / SdfExprEx build, SdfKit/SdfExpr.cs:16-201)
// to the body of `sk_float4 sdf_eval(sk_float3 p)` in the SDF source dialect (csrc/sdfk_prelude.h), i.e. the text
// sdfk_sdf_compile() hands to NVRTC.  It is the C# twin of sdfkit_b200/exprs.py: the tree is evaluated symbolically,
// every System.Single operation becomes one SSA statement, closure captures and parameter-free sub-trees are
// evaluated NOW and printed as exact hexadecimal float literals; any node outside the accepted set
// (SURVEY.md appendix C) throws NotSupportedException -- no CPU fallback.
// SOURCE ONLY: there is no .NET toolchain in this image; this file is not compiled or tested here.
using System;
using System.Collections.Generic;
using System.Globalization;
using System.Linq;
using System.Linq.Expressions;
using System.Numerics;
using System.Reflection;
using System.Text;

namespace SdfKit.B200
{
    public static class SdfExprLowering
    {
        // ---- symbolic values -------------------------------------------------------------------------------
        abstract class Val { }
        sealed class F : Val { public string Name = ""; public float? Const; }             // System.Single
        sealed class B : Val { public string Name = ""; }                                  // comparison result
        sealed class V3 : Val { public F X = null!, Y = null!, Z = null!; }
        sealed class V4 : Val { public F X = null!, Y = null!, Z = null!, W = null!; }
        sealed class Idx : Val { public V3 Position = null!, Index = null!; }              // SdfIndexedInput
        sealed class Obj : Val { public object? Value; }                                   // evaluated closure object

        sealed class Emitter
        {
            public readonly StringBuilder Text = new();
            readonly Dictionary<string, F> cse = new();
            int next;

            public static string Hex(float v)
            {
                if (float.IsNaN(v) || float.IsInfinity(v))
                    return $"sk_bits(0x{BitConverter.SingleToInt32Bits(v):x8}u)";
                if (v == 0) return (1 / v < 0) ? "-0x0.0p+0f" : "0x0.0p+0f";
                // exact binary32 as a C hex-float literal
                int bits = BitConverter.SingleToInt32Bits(v);
                bool neg = bits < 0;
                int exp = (bits >> 23) & 0xFF, man = bits & 0x7FFFFF;
                if (exp == 0) { exp = 1; while ((man & 0x800000) == 0) { man <<= 1; exp--; } man &= 0x7FFFFF; }
                return $"{(neg ? "-" : "")}0x1.{(man << 1):x6}p{(exp - 127 >= 0 ? "+" : "")}{exp - 127}f";
            }

            public F Const(float v) => new F { Name = Hex(v), Const = v };

            public F Op(string fmt, Func<float[], float>? fold, params F[] a)
            {
                if (fold != null && a.All(x => x.Const.HasValue))
                    return Const(fold(a.Select(x => x.Const!.Value).ToArray()));
                var rhs = string.Format(CultureInfo.InvariantCulture, fmt, a.Select(x => (object)x.Name).ToArray());
                if (cse.TryGetValue(rhs, out var hit)) return hit;
                var f = new F { Name = $"t{next++}" };
                Text.Append($"    const float {f.Name} = {rhs};\n");
                return cse[rhs] = f;
            }

            public B Cmp(string op, F a, F b)
            {
                var r = new B { Name = $"c{next++}" };
                Text.Append($"    const bool {r.Name} = {a.Name} {op} {b.Name};\n");
                return r;
            }

            public F Sel(B c, F a, F b) => a.Name == b.Name ? a : Op("sk_sel(" + c.Name + ", {0}, {1})", null, a, b);
        }

        // ---- entry point -----------------------------------------------------------------------------------
        public static string Lower(Expression<SdfFunc> expr)
        {
            var em = new Emitter();
            var p = new V3 { X = new F { Name = "p.x" }, Y = new F { Name = "p.y" }, Z = new F { Name = "p.z" } };
            var env = new Dictionary<ParameterExpression, Val> { [expr.Parameters[0]] = p };
            var r = (V4)Eval(expr.Body, env, em);
            em.Text.Append($"    return sk_make4({r.X.Name}, {r.Y.Name}, {r.Z.Name}, {r.W.Name});\n");
            return em.Text.ToString();
        }

        // ---- helpers ---------------------------------------------------------------------------------------
        static F Add(Emitter e, F a, F b) => e.Op("{0} + {1}", v => v[0] + v[1], a, b);
        static F Sub(Emitter e, F a, F b) => e.Op("{0} - {1}", v => v[0] - v[1], a, b);
        static F Mul(Emitter e, F a, F b) => e.Op("{0} * {1}", v => v[0] * v[1], a, b);
        static F Div(Emitter e, F a, F b) => e.Op("{0} / {1}", v => v[0] / v[1], a, b);
        static F Neg(Emitter e, F a) => e.Op("-({0})", v => -v[0], a);
        static F Call1(Emitter e, string fn, Func<float, float> f, F a) => e.Op(fn + "({0})", v => f(v[0]), a);
        static F Call2(Emitter e, string fn, Func<float, float, float> f, F a, F b) => e.Op(fn + "({0}, {1})", v => f(v[0], v[1]), a, b);
        static V3 Map(Emitter e, V3 a, V3 b, Func<Emitter, F, F, F> f) => new V3 { X = f(e, a.X, b.X), Y = f(e, a.Y, b.Y), Z = f(e, a.Z, b.Z) };
        static V3 Splat(F s) => new V3 { X = s, Y = s, Z = s };
        static F Length(Emitter e, V3 v) =>   // Vector3.Length = sqrt(Dot(v, v)), (xx + yy) + zz
            Call1(e, "sk_sqrt", MathF.Sqrt, Add(e, Add(e, Mul(e, v.X, v.X), Mul(e, v.Y, v.Y)), Mul(e, v.Z, v.Z)));

        static Val Lift(object? o, Emitter e) => o switch
        {
            float f => e.Const(f),
            int i => e.Const(i),
            double d => e.Const((float)d),
            Vector3 v => new V3 { X = e.Const(v.X), Y = e.Const(v.Y), Z = e.Const(v.Z) },
            Vector4 v => new V4 { X = e.Const(v.X), Y = e.Const(v.Y), Z = e.Const(v.Z), W = e.Const(v.W) },
            _ => new Obj { Value = o },
        };

        static bool HasParameters(Expression x)
        {
            var f = new ParamFinder();
            f.Visit(x);
            return f.Found;
        }
        sealed class ParamFinder : ExpressionVisitor
        {
            public bool Found;
            protected override Expression VisitParameter(ParameterExpression node) { Found = true; return node; }
        }

        // ---- the evaluator ---------------------------------------------------------------------------------
        static Val Eval(Expression x, Dictionary<ParameterExpression, Val> env, Emitter e)
        {
            // closure captures / constant sub-trees: evaluate at ToSdf() time
            if (x.NodeType != ExpressionType.Lambda && x.NodeType != ExpressionType.Quote && !HasParameters(x))
                return Lift(Expression.Lambda(x).Compile().DynamicInvoke(), e);

            switch (x)
            {
                case ParameterExpression p:
                    return env.TryGetValue(p, out var v) ? v : throw Unsupported(x);
                case ConstantExpression c:
                    return Lift(c.Value, e);
                case UnaryExpression u when u.NodeType == ExpressionType.Convert:
                    return Eval(u.Operand, env, e);
                case UnaryExpression u when u.NodeType == ExpressionType.Negate:
                    return Eval(u.Operand, env, e) switch
                    {
                        F f => Neg(e, f),
                        V3 v => new V3 { X = Neg(e, v.X), Y = Neg(e, v.Y), Z = Neg(e, v.Z) },
                        _ => throw Unsupported(x),
                    };
                case UnaryExpression u when u.NodeType == ExpressionType.Quote:
                    return new Obj { Value = u.Operand };
                case LambdaExpression l:
                    return new Obj { Value = l };
                case InvocationExpression inv:
                {
                    var target = inv.Expression is LambdaExpression le ? le : (Eval(inv.Expression, env, e) as Obj)?.Value as LambdaExpression;
                    if (target is null) throw Unsupported(x);   // an opaque delegate: cannot be inlined
                    var inner = new Dictionary<ParameterExpression, Val>(env);
                    for (int i = 0; i < target.Parameters.Count; i++) inner[target.Parameters[i]] = Eval(inv.Arguments[i], env, e);
                    return Eval(target.Body, inner, e);
                }
                case BlockExpression blk:
                {
                    var inner = new Dictionary<ParameterExpression, Val>(env);
                    Val last = new Obj();
                    foreach (var s in blk.Expressions)
                    {
                        if (s is BinaryExpression a && a.NodeType == ExpressionType.Assign && a.Left is ParameterExpression lhs)
                            last = inner[lhs] = Eval(a.Right, inner, e);
                        else
                            last = Eval(s, inner, e);
                    }
                    return last;
                }
                case ConditionalExpression cond:
                {
                    var c = Eval(cond.Test, env, e) as B ?? throw Unsupported(x);
                    var a = Eval(cond.IfTrue, env, e); var b = Eval(cond.IfFalse, env, e);
                    return (a, b) switch
                    {
                        (F fa, F fb) => e.Sel(c, fa, fb),
                        (V4 va, V4 vb) => new V4 { X = e.Sel(c, va.X, vb.X), Y = e.Sel(c, va.Y, vb.Y), Z = e.Sel(c, va.Z, vb.Z), W = e.Sel(c, va.W, vb.W) },
                        (V3 va, V3 vb) => new V3 { X = e.Sel(c, va.X, vb.X), Y = e.Sel(c, va.Y, vb.Y), Z = e.Sel(c, va.Z, vb.Z) },
                        _ => throw Unsupported(x),
                    };
                }
                case BinaryExpression bin:
                    return EvalBinary(bin, env, e);
                case MemberExpression m:
                    return EvalMember(m, env, e);
                case NewExpression n:
                    return EvalNew(n, env, e);
                case MemberInitExpression mi when mi.Type == typeof(SdfIndexedInput):
                {
                    var r = new Idx();
                    foreach (var b in mi.Bindings.Cast<MemberAssignment>())
                    {
                        var v = (V3)Eval(b.Expression, env, e);
                        if (b.Member.Name == nameof(SdfIndexedInput.Position)) r.Position = v; else r.Index = v;
                    }
                    return r;
                }
                case MethodCallExpression call:
                    return EvalCall(call, env, e);
            }
            throw Unsupported(x);
        }

        static Val EvalBinary(BinaryExpression b, Dictionary<ParameterExpression, Val> env, Emitter e)
        {
            var l = Eval(b.Left, env, e); var r = Eval(b.Right, env, e);
            if (b.Method != null && b.Method.DeclaringType == typeof(Vector3))
            {
                // Vector3 operator overloads: op_Addition, op_Subtraction, op_Multiply (V3*V3, V3*float, float*V3), op_Division
                var lv = l as V3 ?? Splat((F)l); var rv = r as V3 ?? Splat((F)r);
                return b.NodeType switch
                {
                    ExpressionType.Add => Map(e, lv, rv, Add),
                    ExpressionType.Subtract => Map(e, lv, rv, Sub),
                    ExpressionType.Multiply => Map(e, lv, rv, Mul),
                    ExpressionType.Divide => Map(e, lv, rv, Div),     // Vector3 / float divides per component on .NET Core
                    _ => throw Unsupported(b),
                };
            }
            if (l is F fl && r is F fr)
                return b.NodeType switch
                {
                    ExpressionType.Add => Add(e, fl, fr),
                    ExpressionType.Subtract => Sub(e, fl, fr),
                    ExpressionType.Multiply => Mul(e, fl, fr),
                    ExpressionType.Divide => Div(e, fl, fr),
                    ExpressionType.LessThan => e.Cmp("<", fl, fr),
                    ExpressionType.GreaterThan => e.Cmp(">", fl, fr),
                    _ => throw Unsupported(b),
                };
            throw Unsupported(b);
        }

        static Val EvalMember(MemberExpression m, Dictionary<ParameterExpression, Val> env, Emitter e)
        {
            if (m.Expression is null)   // static: Vector3.One / Zero / UnitX...
                return Lift(m.Member is PropertyInfo pi ? pi.GetValue(null) : ((FieldInfo)m.Member).GetValue(null), e);
            var o = Eval(m.Expression, env, e);
            return (o, m.Member.Name) switch
            {
                (V3 v, "X") => v.X, (V3 v, "Y") => v.Y, (V3 v, "Z") => v.Z,
                (V4 v, "X") => v.X, (V4 v, "Y") => v.Y, (V4 v, "Z") => v.Z, (V4 v, "W") => v.W,
                (Idx i, nameof(SdfIndexedInput.Position)) => i.Position,
                (Idx i, nameof(SdfIndexedInput.Index)) => i.Index,
                _ => throw Unsupported(m),
            };
        }

        static Val EvalNew(NewExpression n, Dictionary<ParameterExpression, Val> env, Emitter e)
        {
            var a = n.Arguments.Select(x => Eval(x, env, e)).ToArray();
            if (n.Type == typeof(Vector3) && a.Length == 3) return new V3 { X = (F)a[0], Y = (F)a[1], Z = (F)a[2] };
            if (n.Type == typeof(Vector3) && a.Length == 1) return Splat((F)a[0]);
            if (n.Type == typeof(Vector4) && a.Length == 4) return new V4 { X = (F)a[0], Y = (F)a[1], Z = (F)a[2], W = (F)a[3] };
            if (n.Type == typeof(Vector4) && a.Length == 2 && a[0] is V3 v) return new V4 { X = v.X, Y = v.Y, Z = v.Z, W = (F)a[1] };
            throw Unsupported(n);
        }

        static Val EvalCall(MethodCallExpression c, Dictionary<ParameterExpression, Val> env, Emitter e)
        {
            var args = c.Arguments.Select(x => Eval(x, env, e)).ToArray();
            var self = c.Object is null ? null : Eval(c.Object, env, e);
            var t = c.Method.DeclaringType;
            string name = c.Method.Name;
            if (t == typeof(Vector3))
            {
                if (name == "Length" && self is V3 sv) return Length(e, sv);
                if (name == "Abs") { var v = (V3)args[0]; return new V3 { X = Call1(e, "sk_abs", MathF.Abs, v.X), Y = Call1(e, "sk_abs", MathF.Abs, v.Y), Z = Call1(e, "sk_abs", MathF.Abs, v.Z) }; }
                if (name == "Max") return Map(e, (V3)args[0], (V3)args[1], (em, p, q) => Call2(em, "sk_vecmax", (x, y) => x > y ? x : y, p, q));
                if (name == "Min") return Map(e, (V3)args[0], (V3)args[1], (em, p, q) => Call2(em, "sk_vecmin", (x, y) => x < y ? x : y, p, q));
            }
            if (t == typeof(MathF) || t == typeof(Math))
            {
                if (name == "Sqrt") return Call1(e, "sk_sqrt", MathF.Sqrt, (F)args[0]);
                if (name == "Abs") return Call1(e, "sk_abs", MathF.Abs, (F)args[0]);
                if (name == "Floor") return Call1(e, "sk_floor", MathF.Floor, (F)args[0]);
                if (name == "Max") return Call2(e, "sk_fmax", MathF.Max, (F)args[0], (F)args[1]);
                if (name == "Min") return Call2(e, "sk_fmin", MathF.Min, (F)args[0], (F)args[1]);
            }
            if (t == typeof(VectorOps))
            {
                if (name == nameof(VectorOps.Mod))   // a - b * floor(a / b)  (SdfKit/VectorData.cs:697-698)
                {
                    var a = (F)args[0]; var b = (F)args[1];
                    return Sub(e, a, Mul(e, b, Call1(e, "sk_floor", MathF.Floor, Div(e, a, b))));
                }
                if (name == nameof(VectorOps.VMax))  // Math.Max(Math.Max(x, y), z)  (SdfKit/VectorData.cs:860-861)
                {
                    var v = (V3)args[0];
                    return Call2(e, "sk_fmax", MathF.Max, Call2(e, "sk_fmax", MathF.Max, v.X, v.Y), v.Z);
                }
            }
            throw Unsupported(c);
        }

        static NotSupportedException Unsupported(Expression x) => new NotSupportedException(
            $"SdfExpr node {x.NodeType} ({x}) cannot be lowered to the GPU dialect; opaque code is rejected, there is no CPU fallback.");
    }
}
