"""Cycle counters of the fused encoder-head kernel's warp roles (CTA 0), from an instrumented build of the library
(-DW2C_HEAD_TIMING -> lib/libw2c_timing.so, built here if missing; the product library carries no instrumentation)."""
import ctypes, sys, torch
import os; sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from multiagentperception_b200 import build as _b
_variant = os.path.join(_b.LIB_DIR, 'libw2c_timing.so')
if not os.path.exists(_variant) or os.environ.get('W2C_REBUILD') == '1':
    _b.build(extra_flags=['-DW2C_HEAD_TIMING'], out=_variant)
os.environ['W2C_LIB'] = _variant
from multiagentperception_b200 import ops, _lib
lib = _lib.load()
dev = torch.device('cuda:0')
b, na, h, w = 8, 5, 512, 512
x = torch.randn(b, 3*na, h, w, device=dev)
w1 = torch.randn(64, 27, device=dev)*0.3; w2 = torch.randn(64, 64, 3, 3, device=dev)/24
s1 = torch.rand(64, device=dev)+0.5; t1 = torch.randn(64, device=dev)*0.1
wp2 = ops.pack_conv_weight(w2, 64, False, 0)
y = ops.new_act(b*na, h//2, w//2, 64, 0, dev)
dbg = torch.zeros(32, dtype=torch.int64, device=dev)
raw = ctypes.CDLL(_variant)
raw.w2c_debug_enc_head_timing.argtypes = [ctypes.c_void_p]
for i in range(3):
    ops.enc_head(x, w1, s1, t1, wp2, s1, t1, y, b=b, n_agents=na, h=h, w=w, act=0)
raw.w2c_debug_enc_head_timing(ctypes.c_void_p(dbg.data_ptr()))
ops.enc_head(x, w1, s1, t1, wp2, s1, t1, y, b=b, n_agents=na, h=h, w=w, act=0)
torch.cuda.synchronize()
d = dbg.cpu().tolist()
tiles = d[13]
print('tiles of CTA0', tiles)
names = {0:'front patch',1:'front wait A1_EMPTY',2:'front im2col',4:'mma wait A1_FULL',5:'mma wait ACC1_EMPTY',6:'mma wait Y1_FULL',7:'mma wait ACC2_EMPTY',8:'epiA wait ACC1_FULL',9:'epiA wait ACC2_FULL',10:'epiA E1',11:'epiA (E2 slot)',12:'epiA total',15:'epiA E1: fence+arrive',23:'epiB fence+arrive',16:'epiB wait ACC1_FULL',17:'epiB wait ACC2_FULL',18:'epiB E1',19:'epiB E2',20:'epiB total'}
for k,v in names.items(): print('%-24s %8.0f clk/tile' % (v, d[k]/max(tiles,1)))
