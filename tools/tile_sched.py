"""List-scheduling bounds from a tile-cost dump (RTC_DUMP_TILE_COST): what the launch would take on `slots` block slots in
natural and in longest-first order, at `mhz`.    python tools/tile_sched.py costs.txt [slots] [mhz]"""
import heapq
import sys

costs = [int(line.split()[-1]) for line in open(sys.argv[1]) if line.strip()]  # lines: band tile cycles
slots = int(sys.argv[2]) if len(sys.argv) > 2 else 148 * 5
mhz = float(sys.argv[3]) if len(sys.argv) > 3 else 1965.0
costs = [c for c in costs if c > 0]


def schedule(order):
    heap = [0] * slots
    heapq.heapify(heap)
    for c in order:
        heapq.heappush(heap, heapq.heappop(heap) + c)
    return max(heap)


total = sum(costs)
srt = sorted(costs)
print(f"{len(costs)} tiles, total {total / 1e6:.1f} M cycles, mean {total / len(costs):.0f}, median {srt[len(srt) // 2]}, "
      f"p99 {srt[int(len(srt) * 0.99)]}, max {srt[-1]}")
print(f"perfect balance on {slots} slots: {total / slots / mhz / 1e3:.4f} ms; longest tile alone: {srt[-1] / mhz / 1e3:.4f} ms")
print(f"list schedule, natural order: {schedule(costs) / mhz / 1e3:.4f} ms; longest first: {schedule(sorted(costs, reverse=True)) / mhz / 1e3:.4f} ms")
