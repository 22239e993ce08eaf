"""Host<->device copy bandwidth on this box: one stream vs several concurrent streams, D2H alone and with H2D running
(the e2e leg of bench.py is PCIe-bound: 31.8 GB per 32768^2 step)."""
import sys
import time

import torch


def run(nbytes, nstreams, direction, with_h2d=False):
    dev = torch.empty(nbytes, dtype=torch.uint8, device="cuda")
    host = torch.empty(nbytes, dtype=torch.uint8).pin_memory()
    streams = [torch.cuda.Stream() for _ in range(nstreams)]
    extra = None
    if with_h2d:
        dev2 = torch.empty(nbytes // 4, dtype=torch.uint8, device="cuda")
        host2 = torch.empty(nbytes // 4, dtype=torch.uint8).pin_memory()
        extra = torch.cuda.Stream()
    chunk = nbytes // nstreams
    best = 0.0
    for _ in range(3):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for k, s in enumerate(streams):
            with torch.cuda.stream(s):
                a, b = k * chunk, (k + 1) * chunk
                if direction == "d2h":
                    host[a:b].copy_(dev[a:b], non_blocking=True)
                else:
                    dev[a:b].copy_(host[a:b], non_blocking=True)
        if extra is not None:
            with torch.cuda.stream(extra):
                dev2.copy_(host2, non_blocking=True)
        for s in streams:
            s.synchronize()
        dt = time.perf_counter() - t0
        best = max(best, nbytes / dt / 1e9)
    return best


def main():
    nbytes = int(float(sys.argv[1]) * 1e9) if len(sys.argv) > 1 else 4 * 10 ** 9
    for direction in ("d2h", "h2d"):
        for ns in (1, 2, 4, 8):
            print("%s %d stream(s): %.1f GB/s" % (direction, ns, run(nbytes, ns, direction)))
    print("d2h 1 stream with h2d running: %.1f GB/s" % run(nbytes, 1, "d2h", True))
    print("d2h 2 streams with h2d running: %.1f GB/s" % run(nbytes, 2, "d2h", True))


if __name__ == "__main__":
    main()
