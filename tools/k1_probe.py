import numpy as np, sys
sys.path.insert(0,'.')
from ros_stereo_slam_b200 import VisualFrontEnd
fe=VisualFrontEnd()
img=(np.random.default_rng(0).integers(0,256,(376,1241))).astype(np.uint8)
try:
    lv,dv=fe.pyramid_level(img,1)
    print("ok", lv.sum())
except Exception as e:
    print("ERR", str(e)[-120:])
