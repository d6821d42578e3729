import os, torch, torch.distributed as dist
if os.environ.get("NCCL_DEBUG", "").upper() == "VERSION":
    os.environ["NCCL_DEBUG"] = "NONE"
local = int(os.environ["LOCAL_RANK"]); torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
t = torch.ones(4, device="cuda"); dist.all_reduce(t); torch.cuda.synchronize()
if dist.get_rank() == 0: print("STDOUT_ONLY_LINE", t[0].item(), "NCCL_DEBUG=", os.environ.get("NCCL_DEBUG"))
dist.destroy_process_group()
