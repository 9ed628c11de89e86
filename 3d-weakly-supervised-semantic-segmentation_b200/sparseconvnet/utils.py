"""scn utility surface used by the reference's drivers (train.py:37,91,94; validation.py:34)."""
import glob
import os

import torch


def is_power2(num):
    return num != 0 and ((num & (num - 1)) == 0)


def _train_state_file(exp_name, name2, epoch):
    return exp_name + "-%09d-" % epoch + name2 + ".train.pth"


def checkpoint_restore(model, exp_name, name2, use_cuda=True, epoch=0, optimizer=None, scheduler=None):
    """Upstream signature and behaviour (train.py:37); `optimizer` / `scheduler` (optional, not in upstream) also restore
    the training state written by checkpoint_save(..., optimizer=, scheduler=) when the file exists -- the reference
    restarts Adam's moments from zero on every resume and rebuilds the scheduler from the epoch alone (train.py:43)."""
    if use_cuda:
        model.cpu()
    if epoch > 0:
        f = exp_name + "-%09d-" % epoch + name2 + ".pth"
        assert os.path.isfile(f)
        print("Restore from " + f)
        model.load_state_dict(torch.load(f))
    else:
        f = sorted(glob.glob(exp_name + "-*-" + name2 + ".pth"))
        f = [x for x in f if not x.endswith(".train.pth")]
        if len(f) > 0:
            f = f[-1]
            print("Restore from " + f)
            model.load_state_dict(torch.load(f))
            epoch = int(f[len(exp_name) + 1:-len(name2) - 5])
    if use_cuda:
        model.cuda()
    if epoch > 0 and (optimizer is not None or scheduler is not None):
        t = _train_state_file(exp_name, name2, epoch)
        if os.path.isfile(t):
            st = torch.load(t)
            if optimizer is not None and st.get("optimizer") is not None:
                optimizer.load_state_dict(st["optimizer"])   # torch casts the moments to each parameter's device
            if scheduler is not None and st.get("scheduler") is not None:
                scheduler.load_state_dict(st["scheduler"])
    return epoch + 1


def checkpoint_save(model, exp_name, name2, epoch, use_cuda=True, optimizer=None, scheduler=None):
    """Upstream signature and behaviour (train.py:91: state_dict via the CPU, previous epoch pruned unless a power of two);
    `optimizer` / `scheduler` (optional) add `<exp>-<epoch>-<name2>.train.pth` with their state_dicts, pruned by the same rule."""
    f = exp_name + "-%09d-" % epoch + name2 + ".pth"
    model.cpu()
    torch.save(model.state_dict(), f)
    if use_cuda:
        model.cuda()
    if optimizer is not None or scheduler is not None:
        torch.save({"optimizer": optimizer.state_dict() if optimizer is not None else None,
                    "scheduler": scheduler.state_dict() if scheduler is not None else None, "epoch": epoch},
                   _train_state_file(exp_name, name2, epoch))
    # remove previous checkpoints unless they are a power of 2 to save disk space
    epoch = epoch - 1
    for f in (exp_name + "-%09d-" % epoch + name2 + ".pth", _train_state_file(exp_name, name2, epoch)):
        if os.path.isfile(f):
            if not is_power2(epoch):
                os.remove(f)
