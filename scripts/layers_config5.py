import sys, ctypes, torch
sys.path.insert(0, '/root/repo')
import demonet_b200
from demonet_b200 import _C, seeded as weights
B, S = (int(sys.argv[1]) if len(sys.argv) > 1 else 512), 512
model = demonet_b200.ssd_lite_mobilenet_v2(image_size=S); model.load_state_dict(weights.seeded_state_dict(model.state_dict())); model = model.cuda()
eng = model.reserve(B); imgs = weights.synthetic_images(B, S).cuda()
per = (ctypes.c_float * (len(model.plan.layers) + 3))()
_C.check(_C.lib().dn_engine_profile(eng._handle, imgs.data_ptr(), B, 5, per, torch.cuda.current_stream().cuda_stream))
for i, (L, t) in enumerate(zip(model.plan.layers, list(per))):
    if L.kind == 'dw': by = B*(L.h_in*L.w_in + L.h_out*L.w_out)*L.cin*2
    elif L.kind == 'pw': by = B*L.h_in*L.w_in*(L.cin*2 + L.cout*(4 if L.head else 2)) + (B*L.h_in*L.w_in*L.cout*2 if L.res else 0)
    else: by = B*(3*S*S*4 + L.h_out*L.w_out*L.cout*2)
    print('%3d %-4s %3dx%-3d c%4d->%4d k%d s%d %8.4f ms %7.1f GB/s' % (i, L.kind, L.h_in, L.w_in, L.cin, L.cout, L.k, L.stride, t, by/(t*1e-3)/1e9))
print('post', list(per)[len(model.plan.layers):])
