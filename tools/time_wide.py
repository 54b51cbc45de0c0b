import sys; sys.path.insert(0,'/root/repo'); sys.path.insert(0,'/root/repo/tests')
import torch, time
from conftest import make_inputs
from wav2sleep_b200 import build_default
EOG={'EOG-L':'EOG-L','EOG-R':'EOG-R'}
model=build_default(EOG,5,seed=0).cuda().eval()
x={k:v.cuda() for k,v in make_inputs(EOG,16,1680,seed=1).items()}
for wide in (0,2,4):
    for enc in model.signal_encoders.encoders.values(): enc.wide_blocks=wide
    for _ in range(2): model(x)
    torch.cuda.synchronize(); t=time.time()
    for _ in range(5): model(x)
    torch.cuda.synchronize(); print('wide',wide,'ms',(time.time()-t)/5*1e3)
