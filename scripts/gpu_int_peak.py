import sys
sys.path.insert(0, ".")
from crispresso_b200 import Context
ctx = Context(0)
names = ["IADD (compiler splits alu/fma)", "IMAD", "VIMNMX.S16x2", "VIADDMNMX.S16x2", "VIMNMX.S16x2 + IMAD 1:1", "VIMNMX3.S16x2", "LOP3", "sub+VIMNMX imm+add", "HSET2.NE + LOP3 (2 ops)", "HSET2.NE"]
for w, n in enumerate(names):
    print("which=%d %-34s %.3f Tlane-op/s" % (w, n, ctx.int_peak(w) / 1e12))
