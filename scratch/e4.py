"""phase stamps of one HeadCmaxPair CTA (REC_CMAX_TRACE): per-unit intervals of issuer / epilogue / loader"""
import sys, numpy as np
t = np.fromfile(sys.argv[1], dtype=np.int64).reshape(8, 64)
names = ["iss:acc", "iss:stage", "iss:issued", "epi0:start", "epi0:end", "ld:free", "epi15:start", "epi15:end"]
t0 = t[t > 0].min()
print("unit " + " ".join(f"{n:>11s}" for n in names))
for i in range(8, 28):
    print(f"{i:4d} " + " ".join(f"{int(t[k, i] - t0):11d}" for k in range(8)))
d = lambda k: np.diff(t[k, 8:60]).mean()
print("period per kind:", [round(float(d(k))) for k in range(8)])
print("issuer: acc->stage", (t[1] - t[0])[8:60].mean(), "stage->issued", (t[2] - t[1])[8:60].mean(), "issued->next acc", (t[0, 9:61] - t[2, 8:60]).mean())
print("epi0: tile", (t[4] - t[3])[8:60].mean(), "end->next start", (t[3, 9:61] - t[4, 8:60]).mean())
print("epi15: tile", (t[7] - t[6])[8:60].mean(), "end->next start", (t[6, 9:61] - t[7, 8:60]).mean())
print("issued(i) -> epi0 start(i)", (t[3] - t[2])[8:60].mean(), " epi0 end(i) -> iss acc(i+2)", (t[0, 10:62] - t[4, 8:60]).mean())
