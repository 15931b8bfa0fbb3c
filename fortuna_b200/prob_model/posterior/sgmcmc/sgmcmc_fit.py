"""The SG-MCMC fit shared by SGHMCPosterior and CyclicalSGLDPosterior - the part of
fortuna/prob_model/posterior/sgmcmc/sghmc/sghmc_posterior.py:62-187 around the sampler: build the integrator,
create the state, run ``trainer.train`` with the log-joint as the (ascended) objective and the sampling callback.

The host loop is the reference's ``_training_loop`` (fortuna/training/trainer.py:455-532) reduced to what it does for
SG-MCMC: per batch, value_and_grad of the batched log-joint (``Joint._batched_log_joint_prob``,
fortuna/prob_model/joint/base.py:54-122: log prior + n_data / batch_size * sum log-likelihood) by autograd, then
``state.apply_gradients`` - ONE sampler-kernel launch - then the callbacks.

With the isotropic Gaussian prior (the only prior of the hot path) autograd sees the batch log-likelihood SUM only: the
``n_data / batch_size`` scaling and the prior gradient ``-prec * theta`` are formed inside the sampler kernel
(fb200_sgmcmc_step_joint_f32), which also returns ``sum(theta^2)`` so the logged log-joint value is complete.  The
unfused route (prior through autograd) stays available as ``fuse_prior=False`` - the two are compared in the tests.

Data-parallel (``bound.dp`` set, one process per GPU): the reference's multi-device trainer gives every device a
contiguous slice of each batch (fortuna/training/trainer.py:700-760), evaluates the log-joint on it with the slice's own
``n_data / batch`` scaling, and averages gradients and loss over the devices (``lax.pmean``, :793-795).  Here every rank
takes its slice, back-propagates the slice's log-likelihood sum into its peer-mapped gradient vector, and ONE kernel per
rank does average + prior + step + parameter broadcast (fb200_sgmcmc_step_dp_f32); the logged loss is averaged over the
ranks once per epoch."""
import math
from typing import Dict, List

import numpy as np
import torch

from ....training.train_state import TrainState
from .torch_model import BoundModel


def categorical_log_likelihood_sum(outputs: torch.Tensor, targets: torch.Tensor) -> torch.Tensor:
    """sum_b log softmax(outputs_b)[y_b] - fortuna/prob_output_layer/classification.py:25-28."""
    return -torch.nn.functional.cross_entropy(outputs, targets.long(), reduction="sum")


def batched_log_joint_prob(bound: BoundModel, prior, inputs: torch.Tensor, targets: torch.Tensor, n_data: int,
                           log_likelihood_sum=categorical_log_likelihood_sum) -> torch.Tensor:
    ll = log_likelihood_sum(bound.model(inputs), targets)
    lp = sum(prior.log_joint_prob(p) for p in bound.trainable.values())
    return lp + (n_data / inputs.shape[0]) * ll


def run_sgmcmc(bound: BoundModel, prior, tx, callbacks: List, train_data_loader, n_epochs: int, val_data_loader=None,
               log_likelihood_sum=categorical_log_likelihood_sum, save_chain=None,
               save_every_n_steps=None, fuse_prior: bool = True, max_grad_norm=None) -> Dict[str, np.ndarray]:
    """``save_chain(state)`` writes the chain checkpoint (``<save_checkpoint_dir>/c``): every ``save_every_n_steps``
    steps (fortuna/training/trainer.py:236-247) and once when training ends (:146-150)."""
    state = TrainState.create(params=bound.params, tx=tx)
    n_data = train_data_loader.size
    dev = bound.device
    dp = getattr(bound, "dp", None)
    if dp is not None and not (fuse_prior and hasattr(prior, "log_var")):
        raise NotImplementedError("the data-parallel fit needs the fused Gaussian-prior route")
    fuse = fuse_prior and hasattr(prior, "log_var")
    prec = math.exp(-prior.log_var) if fuse else None
    lp_const = -0.5 * bound.params.n_params * (math.log(2 * math.pi) + prior.log_var) if fuse else None
    if dp is not None and save_chain is not None:
        # every rank holds only its owned slice of the momentum / preconditioner moments: gather them, then ONE rank
        # writes (all ranks writing `<dir>/c/checkpoint_tmp` and renaming it raced, and each file held mostly zeros)
        write_chain = save_chain

        def save_chain(st):  # noqa: F811
            if tx.dp_gather is not None:
                tx.dp_gather(st.opt_state, dp)
            if dp.rank == 0:
                write_chain(st)

    losses, val_losses = [], []
    for _ in range(n_epochs):
        epoch = []
        for inputs, targets in train_data_loader:
            x = torch.as_tensor(inputs, dtype=torch.float32, device=dev)
            y = torch.as_tensor(targets, device=dev)
            bound.zero_grad()
            if dp is not None:
                if x.shape[0] % dp.world:
                    raise ValueError(f"the batch size ({x.shape[0]}) must be a multiple of the number of devices ({dp.world})")
                b = x.shape[0] // dp.world
                x, y = x[dp.rank * b : (dp.rank + 1) * b], y[dp.rank * b : (dp.rank + 1) * b]
                ll = log_likelihood_sum(bound.model(x), y)
                ll.backward()
                scale = n_data / b
                # the log prior of the pre-update parameters, for the logged value only (the step kernel forms its gradient)
                with torch.no_grad():
                    lp = sum(prior.log_joint_prob(p) for p in bound.trainable.values())
                state = state.apply_gradients(grads=bound.grads, joint=(scale, prec, None), dp=dp, max_grad_norm=max_grad_norm)
                lj = (lp.double() + scale * ll.detach().double()).float()
            elif fuse:
                ll = log_likelihood_sum(bound.model(x), y)
                ll.backward()                                # d sum-log-lik / d theta lands in bound.grads (flat)
                scale = n_data / x.shape[0]
                sq = torch.empty(1, dtype=torch.float64, device=dev)
                state = state.apply_gradients(grads=bound.grads, joint=(scale, prec, sq), max_grad_norm=max_grad_norm)
                lj = (lp_const - 0.5 * prec * sq[0] + scale * ll.detach().double()).float()
            else:
                lj = batched_log_joint_prob(bound, prior, x, y, n_data, log_likelihood_sum)
                lj.backward()                                    # d log-joint / d theta lands in bound.grads (flat)
                state = state.apply_gradients(grads=bound.grads, max_grad_norm=max_grad_norm)  # fused sampler step, in place
            for cb in callbacks:
                state = cb.training_step_end(state)
            if save_chain is not None and save_every_n_steps and state.step % save_every_n_steps == 0:
                save_chain(state)
            epoch.append(lj.detach())
        ep_mean = torch.stack(epoch).mean()
        if dp is not None:  # pmean of the loss (trainer.py:793-795), once per epoch
            import torch.distributed as dist

            dist.all_reduce(ep_mean)
            ep_mean = ep_mean / dp.world
        losses.append(float(ep_mean))
        if val_data_loader is not None:
            with torch.no_grad():
                v = [batched_log_joint_prob(bound, prior, torch.as_tensor(a, dtype=torch.float32, device=dev),
                                            torch.as_tensor(b, device=dev), val_data_loader.size, log_likelihood_sum)
                     for a, b in val_data_loader]
            val_losses.append(float(torch.stack(v).mean()))
    if save_chain is not None:
        save_chain(state)
    status = {"loss": np.array(losses)}
    if val_losses:
        status["val_loss"] = np.array(val_losses)
    return status, state
