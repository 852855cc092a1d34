"""Generate the numerical constants of the deterministic-math spec.

Run once; the printed hex-float literals are pasted into oracle/sabc_oracle.c and
simulatedannealingabc.jl_b200/csrc/detmath.cuh (two independent restatements of one spec).
"""
import mpmath as mp
mp.mp.prec = 200

def hx(x):
    return float(x).hex()

print("// sin(pi r) = r * (S0 + r2*(S1 + ...)) ; S_k = (-1)^k pi^(2k+1)/(2k+1)!")
for k in range(10):
    c = (-1) ** k * mp.pi ** (2 * k + 1) / mp.factorial(2 * k + 1)
    print(f"S{k} = {hx(c)}  /* {mp.nstr(c, 20)} */")
print("// cos(pi r) = C0 + r2*(C1 + ...) ; C_k = (-1)^k pi^(2k)/(2k)!")
for k in range(10):
    c = (-1) ** k * mp.pi ** (2 * k) / mp.factorial(2 * k)
    print(f"C{k} = {hx(c)}  /* {mp.nstr(c, 20)} */")
print("// log(k!) k=0..16")
for k in range(17):
    c = mp.log(mp.factorial(k))
    print(f"LF{k} = {hx(c)}  /* {mp.nstr(c, 20)} */")
print("LOG2PI =", hx(mp.log(2 * mp.pi)), mp.nstr(mp.log(2 * mp.pi), 20))
print("HALF_LOG2PI =", hx(mp.log(2 * mp.pi) / 2), mp.nstr(mp.log(2 * mp.pi) / 2, 20))
print("LN2_HI =", float.fromhex('0x1.62e42fee00000p-1'), "LN2_LO =", hx(mp.log(2) - mp.mpf(float.fromhex('0x1.62e42fee00000p-1'))))
print("INV_LN2 =", hx(1 / mp.log(2)))
