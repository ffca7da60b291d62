import random, subprocess, sys
P = {'q': 0x30644e72e131a029b85045b68181585d97816a916871ca8d3c208c16d87cfd47,
     'r': 0x30644e72e131a029b85045b68181585d2833e84879b9709143e1f593f0000001}
R = 1 << 256
random.seed(7)
def h(x): return "%064x" % x
cases = []
def edge(m):
    return random.choice([0, 1, 2, m-1, m-2, (m-1)//2, 1<<255 if (1<<255)<m else m-3, (1<<52)-1, 1<<52, (1<<51), (1<<104)-1, m - (1<<51), random.getrandbits(60)])
def rnd(m, k):
    return edge(m) if random.random() < 0.3 else random.randrange(m)
for f in 'qr':
    m = P[f]
    Ri = pow(R, -1, m)
    for k in range(1500):
        a, b, c, d = (rnd(m, k) for _ in range(4))
        # std-montgomery semantics: operands X = x*R; results mont products
        cases.append((f, 'mul', [a, b], a*b*Ri % m))
        cases.append((f, 'sqr', [a], a*a*Ri % m))
        cases.append((f, 'dual', [a, b, c, d], (a*b + c*d)*Ri % m))
        cases.append((f, 'conv', [a], a % m))
        cases.append((f, 'submul', [a, b, c], (a-b)*c*Ri % m))
        cases.append((f, 'lazy', [a, b, c], (a-b)*c*Ri % m))
        cases.append((f, 'lazy3', [a, b, c, d], (a-b-2*c)*d*Ri % m))
        x = a*b*Ri % m; x = x*x*Ri % m; x = (x - c) % m; x = x*d*Ri % m; x = x*x*Ri % m
        cases.append((f, 'chain', [a, b, c, d], x))
inp = "\n".join("%s %s %s" % (f, op, " ".join(h(x) for x in args)) for f, op, args, _ in cases) + "\n"
out = subprocess.run([sys.argv[1]], input=inp, capture_output=True, text=True, check=True).stdout.split("\n")
bad = 0; mx = 0
for (f, op, args, exp), line in zip(cases, out):
    got, ml = line.split()
    mx = max(mx, float(ml))
    if int(got, 16) != exp:
        bad += 1
        if bad < 10: print("MISMATCH", f, op, [h(a) for a in args], got, h(exp))
print("cases", len(cases), "bad", bad, "max log2 limb", mx)
