import numpy as np, sys
d = np.loadtxt(sys.argv[1], dtype=np.int64)
c = d[:,2].astype(float)
print("tiles", len(c), "sum", c.sum(), "mean", c.mean(), "max", c.max(), "p99", np.percentile(c,99), "p90", np.percentile(c,90), "median", np.median(c))
slots = 148*5
print("total/slots", c.sum()/slots, " max/ (total/slots)", c.max()/(c.sum()/slots))
# simulate LPT list scheduling with `slots` machines
import heapq
h=[0.0]*slots
for x in c: 
    t=heapq.heappop(h); heapq.heappush(h,t+x)
print("LPT makespan", max(h), "natural-order sim:")
