"""Size-independent checks of the largest single-GPU transforms (2^25..2^27): inverse(forward(x)) == x, X[0] == sum x,
and the coset LDE 2^20 -> 2^25 round trip."""
import sys
import torch
sys.path.insert(0, ".")
from toyni_b200 import device as D
from toyni_b200.lib import P
ok = True
for log_n in (25, 26, 27):
    g = torch.Generator(device="cuda"); g.manual_seed(log_n)
    x = torch.randint(0, P, (1 << log_n,), dtype=torch.int32, device="cuda", generator=g)
    f = D.ntt_(x.clone())
    good = int(f[0]) == int(x.to(torch.int64).sum() % P) and torch.equal(D.ntt_(f, True), x)
    good &= int(f.max()) < P and int(f.min()) >= 0
    print(f"2^{log_n}: {'OK' if good else 'FAIL'}", flush=True)
    ok &= bool(good)
    del x, f
c = torch.randint(0, P, (1 << 20,), dtype=torch.int32, device="cuda")
back = D.coset_ifft_(D.coset_fft(c, 1 << 25, 7), 7)
good = torch.equal(back[: 1 << 20], c) and int(back[1 << 20:].abs().max()) == 0
print("LDE 2^20 -> 2^25 round trip:", "OK" if good else "FAIL")
sys.exit(0 if ok and good else 1)
