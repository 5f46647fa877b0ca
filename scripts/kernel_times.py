import sys, os, numpy as np, torch
sys.path.insert(0, '/root/repo')
from tezip_b200 import synth, ops
from tezip_b200.prednet import PredNet
STACK=(3,48,96,192)
ws=synth.make_weights(STACK,bias="uniform",seed=7)
net=PredNet(STACK,STACK,weights=ws,input_hw=(128,160),max_batch=100)
fr=torch.from_numpy(synth.make_frames(100,128,160,3,seed=1)).cuda()
x=ops.pad_normalize(fr,None,128,160); out=torch.empty_like(x)
acc=np.zeros(len(net.kernels()))
for i in range(8):
    ms=net.next_timed(x,out)
    if i>=2: acc+=np.array(ms)/6
print(os.environ.get("TAG",""), "sum %.4f"%acc.sum(), " ".join("%s=%.4f"%(n[0].replace("conv_tc_",""),m) for n,m in zip(net.kernels(),acc)))
# whole chain of 9 steps, untimed events inside: PDL active
o2=torch.empty_like(x)
def chain():
    net.next(x,out=out)
    a,b=out,o2
    for k in range(8):
        net.next_chained(b); a,b=b,a
for _ in range(3): chain()
torch.cuda.synchronize()
e0,e1=torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(5): chain()
e1.record(); torch.cuda.synchronize()
print(os.environ.get("TAG",""), "9-step chain ms %.4f"%(e0.elapsed_time(e1)/5), "checksum", float(out.double().sum()))
