"""Entry point with the behaviour of GA3C/Run.py:74 (`Server().main()`): `python -m rl_collision_avoidance_b200.ga3c.Run`.
The config class comes from GYM_CONFIG_CLASS / GYM_CONFIG_PATH exactly as train.sh:9-15 exports them.  Under torchrun
(one rank per GPU) worlds are sharded across ranks."""
import os


def main():
    import torch
    if "RANK" in os.environ and int(os.environ.get("WORLD_SIZE", "1")) > 1:
        torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", "0")))
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        import datetime
        torch.distributed.init_process_group("nccl", timeout=datetime.timedelta(seconds=int(os.environ.get("GA3C_NCCL_TIMEOUT_S", "600"))))
    from .Server import Server
    out = Server().main(max_seconds=float(os.environ["GA3C_MAX_SECONDS"]) if "GA3C_MAX_SECONDS" in os.environ else None)
    print("all done.", out)
    if torch.distributed.is_available() and torch.distributed.is_initialized():
        torch.distributed.destroy_process_group()


if __name__ == "__main__":
    main()
