import sys, numpy as np, torch
sys.path.insert(0, '/root/repo')
from oracle import geodesy
from vissatsatellitestereo_b200 import engine as E, synthetic as S
cfg = S.scaled(S.CONFIGS['C1'], views=1, depth=64, grid=32)
eng = E.DsmEngine(S.make_aoi(cfg, geodesy), cfg.res, cfg.res)
img = torch.randn(96, 96, device='cuda')
out = eng.median3x3(img)
torch.cuda.synchronize()
import cv2
print('ok', np.array_equal(out.cpu().numpy(), cv2.medianBlur(img.cpu().numpy(), 3)))
