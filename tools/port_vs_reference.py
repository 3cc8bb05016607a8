import sys, time, os
sys.path.insert(0,'/root/repo'); sys.path.insert(0,'/root/repo/comprehensive-transformer-tts_b200'); sys.path.insert(0,'/root/repo/tests/golden')
import torch
import bench
from oracle import ctts_oracle as O
from oracle.ref_import import import_reference, reference_configs
torch.set_num_threads(8)
(p,m,t), sd, batch, frames = bench.build_workload(0)
a = (batch["speakers"], batch["texts"], batch["src_lens"], batch["max_src_len"])
def timeit(fn, n=3):
    fn(); ts=[]
    for _ in range(n):
        t0=time.perf_counter(); fn(); ts.append(time.perf_counter()-t0)
    return sorted(ts)[len(ts)//2]
with torch.no_grad():
    tp = timeit(lambda: O.comp_trans_tts_forward(sd,p,m,t,*a))
cwd=os.getcwd()
ref_model,_ = import_reference()
rp,rm,rt = reference_configs("LJSpeech")
rm["block_type"]="transformer_fs2"; rm["duration_modeling"]["learn_alignment"]=False
net = ref_model.CompTransTTS(rp,rm,rt); net.load_state_dict(sd, strict=True); net.eval()
os.chdir(cwd)
with torch.no_grad():
    tr = timeit(lambda: net(*a))
print("port %.0f ms, reference %.0f ms, ratio %.3f" % (tp*1e3, tr*1e3, tp/tr))
# profile port
import torch.autograd.profiler as prof
with torch.no_grad(), prof.profile() as pr:
    O.comp_trans_tts_forward(sd,p,m,t,*a)
print(pr.key_averages().table(sort_by="self_cpu_time_total", row_limit=12))
with torch.no_grad(), prof.profile() as pr2:
    net(*a)
print(pr2.key_averages().table(sort_by="self_cpu_time_total", row_limit=12))
