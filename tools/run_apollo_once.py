"""One Apollo restorer forward + one MDX stft / istft pair inside a cudaProfilerStart/Stop range (ncu target):

  ncu --metrics <list of tools/ncu_steps.sh> --clock-control none --profile-from-start off -o gpurun_out/apollo \
      python tools/run_apollo_once.py [seconds = 10]
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
from targetdiarization_b200 import ConvTDFNet, Restorer, synth  # noqa: E402

secs = float(sys.argv[1]) if len(sys.argv) > 1 else 10.0
ns = int(secs * 44100)
rest = Restorer(synth.random_apollo_state_dict(0), "cuda:0")
x = synth.synthetic_fullband(1, ns, seed=1).reshape(1, 1, ns).cuda()
net = ConvTDFNet("vocals", 11, 3072, 8, 6144, 1024, "cuda:0")
xm = torch.randn(4, 2, net.chunk_size, device="cuda") * 0.1
rest(x)
spec = net.stft(xm)
net.istft(spec)
torch.cuda.synchronize()
torch.cuda.profiler.start()
rest(x)
spec = net.stft(xm)
net.istft(spec)
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print("apollo forward of", secs, "s + mdx stft/istft of 4 stereo chunks")
