"""CPU: host-side logic of the fused optimizer that needs no device — whole-object pickling (the reference's optimizer
checkpoint format, train_encoder.py:209,217,413: ``torch.save(optimizer)`` / ``optimizer = torch.load(...)``) and the
muP parameter-group split."""
import io

import torch


def test_fused_adamw_survives_whole_object_pickle():
    from omnibiote_b200.optim import FusedAdamW
    p = torch.nn.Parameter(torch.zeros(8, 8, dtype=torch.bfloat16))
    opt = FusedAdamW([p], lr=1e-3, weight_decay=1e-2)
    opt.state[p]["step"] = torch.tensor(7.0)
    opt.state[p]["exp_avg"] = torch.ones_like(p)
    opt.state[p]["exp_avg_sq"] = torch.ones_like(p)
    opt._plan_key, opt._plan, opt._global_step = ("stale",), {"stale": True}, 7
    buf = io.BytesIO()
    torch.save(opt, buf)
    buf.seek(0)
    o2 = torch.load(buf, weights_only=False)
    # torch's Optimizer.__getstate__ keeps defaults / state / param_groups only: the launch plan must come back empty
    # (rebuilt lazily by the next clip_and_step) instead of missing, and the step count follows the saved state
    assert o2._plan_key is None and o2._plan is None and o2.last_grad_norm is None
    assert o2._global_step == 7
    (q,) = o2.param_groups[0]["params"]
    assert float(o2.state[q]["step"]) == 7.0 and torch.equal(o2.state[q]["exp_avg"], torch.ones_like(q))
    assert o2.param_groups[0]["lr"] == 1e-3 and o2.param_groups[0]["weight_decay"] == 1e-2


def test_state_dict_round_trip_keeps_reference_keys():
    from omnibiote_b200.optim import FusedAdamW
    p = torch.nn.Parameter(torch.zeros(4, 4, dtype=torch.bfloat16))
    opt = FusedAdamW([p], lr=1e-3)
    opt.state[p]["step"] = torch.tensor(3.0)
    opt.state[p]["exp_avg"] = torch.full_like(p, 0.5)
    opt.state[p]["exp_avg_sq"] = torch.full_like(p, 0.25)
    sd = opt.state_dict()
    assert set(sd["state"][0].keys()) == {"step", "exp_avg", "exp_avg_sq"}   # torch.optim.AdamW's keys
    q = torch.nn.Parameter(torch.zeros(4, 4, dtype=torch.bfloat16))
    o2 = FusedAdamW([q], lr=5e-4)
    o2.load_state_dict(sd)
    assert torch.equal(o2.state[q]["exp_avg"], torch.full_like(q, 0.5)) and o2.param_groups[0]["lr"] == 1e-3


def test_scheduler_sees_the_optimizer_step():
    """LinearLR wraps optimizer.step to detect `scheduler.step()` before `optimizer.step()`; the trainer's fused
    clip + step goes through step(), so the bookkeeping flag must be set by it (no UserWarning every run)."""
    from omnibiote_b200.optim import FusedAdamW
    p = torch.nn.Parameter(torch.zeros(4, 4, dtype=torch.bfloat16))
    opt = FusedAdamW([p], lr=1e-3)
    sched = torch.optim.lr_scheduler.LinearLR(opt, start_factor=1.0, end_factor=0.0, total_iters=10)
    assert hasattr(opt.step, "_wrapped_by_lr_sched")
    opt.step(max_norm=1.0, grad_scale=1.0, zero_grad=True)   # no gradients yet: nothing to launch, flag still set
    assert getattr(opt, "_opt_called", False)
    sched.step()
    assert abs(opt.param_groups[0]["lr"] - 0.9e-3) < 1e-12
