"""Where does a step's time go on the host?  Prints per-step CPU call time, caching-allocator
device-malloc counts and GPU time for a module at a BASELINE shape."""
import sys, time, os
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import nnaudio_b200 as nb

which = sys.argv[1] if len(sys.argv) > 1 else "stft"
if which == "stft":
    mod = nb.STFT(n_fft=2048, hop_length=512, output_format="Magnitude", verbose=False).cuda()
else:
    mod = nb.MelSpectrogram(sr=22050, n_fft=2048, hop_length=512, n_mels=128, verbose=False).cuda()
xs = [torch.randn(64, 220500, device="cuda") for _ in range(3)]
with torch.no_grad():
    for i in range(5):
        y = mod(xs[i % 3])
    torch.cuda.synchronize()
    st0 = torch.cuda.memory_stats()
    cpu = []
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    t_all = time.perf_counter()
    for i in range(20):
        t0 = time.perf_counter()
        y = mod(xs[i % 3])
        cpu.append(time.perf_counter() - t0)
    ev1.record()
    torch.cuda.synchronize()
    wall = time.perf_counter() - t_all
    st1 = torch.cuda.memory_stats()
print(which, "gpu ms/step", ev0.elapsed_time(ev1) / 20, "wall ms/step", wall * 50,
      "cpu call ms (median)", sorted(cpu)[10] * 1e3, "max", max(cpu) * 1e3)
for k in ("num_device_alloc", "num_device_free", "num_alloc_retries", "allocation.all.allocated"):
    print("  ", k, st1.get(k, 0) - st0.get(k, 0))
print("   reserved MB", torch.cuda.memory_reserved() / 1e6, "allocated MB", torch.cuda.memory_allocated() / 1e6)
