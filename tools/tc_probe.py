"""GPU probe for the tcgen05 layer kernel: one layer, both modes, both descriptor conventions."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from ev2hands_b200 import _capi

torch.manual_seed(0)
dev = "cuda:0"
for (M, cin, cout, pool) in [(128, 32, 64, 0), (128, 32, 16, 0), (1000, 8, 32, 0), (4096, 64, 96, 0), (4096, 323, 256, 0),
                             (2048, 128, 256, 128), (2048, 64, 128, 64), (2048, 32, 64, 32), (1024, 515, 512, 0)]:
    ld = (cin + 3) // 4 * 4
    x = torch.zeros(M, ld, device=dev); x[:, :cin] = torch.randn(M, cin, device=dev)
    w = torch.randn(cout, cin, device=dev) / cin ** 0.5
    b = torch.randn(cout, device=dev) * 0.1
    one, zero = torch.ones(cout, device=dev), torch.zeros(cout, device=dev)
    wt, bias = _capi.fold_conv_bn(w, b, one, zero, zero, one - 1e-5, 1e-5)
    want = torch.relu(x[:, :cin].double() @ w.double().t() + b.double())
    if pool:
        want = want.view(-1, pool, cout).max(1).values
    for mode, name in ((1, "tf32x3"), (0, "bf16")):
        for dbg in (0,):
            _capi.lib().ev2h_tc_set_debug(dbg)
            rows = M // pool if pool else M
            y = torch.zeros(rows, cout, device=dev)
            packed = _capi.tc_pack(wt, cin, cout, mode)
            try:
                _capi.linear_relu_tc(x, M, ld, cin, packed, bias, cout, pool, y, cout, 0, mode)
                torch.cuda.synchronize()
                err = ((y.double() - want).abs().max() / want.abs().max()).item()
            except Exception as ex:
                err = repr(ex)[:100]
            print("M=%d cin=%d cout=%d pool=%d %s swap=%d  rel_err=%s" % (M, cin, cout, pool, name, dbg, err), flush=True)
_capi.lib().ev2h_tc_set_debug(0)
