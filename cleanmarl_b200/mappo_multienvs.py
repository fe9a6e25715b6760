"""Drop-in for ``cleanmarl/mappo_multienvs.py``: same ``Args`` / tyro CLI (MME:18-79, 289), same run
directory and TensorBoard tags (MME:345-362, 461-465, 605-612, 642-644), with the whole training path on
the GPU.

    python cleanmarl_b200/mappo_multienvs.py --env_type pz --env_name simple_spread_v3 --batch_size 4096
    torchrun --nproc-per-node 8 cleanmarl_b200/mappo_multienvs.py --batch_size 65536      # envs sharded, one
                                                                                          # all-reduce per epoch
"""
from __future__ import annotations

import datetime
import sys
from pathlib import Path

if __package__ in (None, ""):                       # executed as a script
    sys.path.insert(0, str(Path(__file__).resolve().parent.parent))

import tyro  # noqa: E402

from cleanmarl_b200.mappo import MAPPO, Args, evaluate, init_distributed, validate_args  # noqa: E402

ALGO = "MAPPO"
IPPO = False


def main(argv=None, algo=ALGO, ippo=IPPO, args_cls=Args, run_prefix=None, trainer_cls=MAPPO, evaluate_fn=evaluate,
         SummaryWriter=None):
    """``trainer_cls`` / ``evaluate_fn`` / ``SummaryWriter`` exist for the CPU tests of this loop (tests/test_host_cpu.py runs
    it on the engine's test double with a recording writer); the scripts never pass them."""
    args = tyro.cli(args_cls, args=argv)
    validate_args(args)
    rank, world, local = init_distributed()
    trainer = trainer_cls(args, device_index=local, rank=rank, world_size=world, ippo=ippo)
    writer = None
    run_prefix = run_prefix or f"{algo}-multienvs"
    if rank == 0:
        if SummaryWriter is None:
            from torch.utils.tensorboard import SummaryWriter
        time_token = datetime.datetime.now().strftime("%Y-%m-%d_%H-%M-%S")
        run_name = f"{args.env_type}__{args.env_name}__{time_token}"
        if args.use_wnb:
            import wandb
            wandb.init(project=args.wnb_project, entity=args.wnb_entity, sync_tensorboard=True, config=vars(args),
                       name=f"{run_prefix}-{run_name}")
        writer = SummaryWriter(f"runs/{run_prefix}-{run_name}")
        writer.add_text("hyperparameters", "|param|value|\n|-|-|\n%s" % (
            "\n".join([f"|{key}|{value}|" for key, value in vars(args).items()])))
    # the reference keeps every episode return since the last log line and logs their mean once MORE than
    # ``log_every`` of them are pending (MME:454-468): the same mean from the per-rollout means
    pending_episodes, pending_reward_sum = 0, 0.0

    def log(record):
        """Scalars of one iteration, in the reference's order (MME:460-468, 605-612, 642-644)."""
        nonlocal pending_episodes, pending_reward_sum
        handle, step, training_step, num_episodes, ev = record
        sc = trainer.read_scalars(handle)                           # the only host synchronisation of the loop
        pending_episodes += args.batch_size
        pending_reward_sum += sc["ep_reward"] * args.batch_size
        if writer is None:
            return
        if pending_episodes > args.log_every:
            writer.add_scalar("rollout/ep_reward", pending_reward_sum / pending_episodes, step)
            writer.add_scalar("rollout/ep_length", sc["ep_length"], step)
            writer.add_scalar("rollout/num_episodes", num_episodes, step)
            pending_episodes, pending_reward_sum = 0, 0.0
        for k in ("actor_loss", "critic_loss", "entropy", "kl_divergence", "clipped_ratios", "actor_gradients",
                  "critic_gradients"):                              # MME:605-612
            writer.add_scalar(f"train/{k}", sc[k], step)
        writer.add_scalar("train/num_updates", training_step, step)
        if ev is not None:
            writer.add_scalar("eval/ep_reward", ev[0], step)
            writer.add_scalar("eval/std_ep_reward", ev[1], step)
            writer.add_scalar("eval/ep_length", ev[2], step)

    # The scalars of iteration k are written AFTER iteration k + 1 has been launched: TensorBoard / W&B I/O on rank 0
    # then overlaps device work instead of keeping the other ranks' gradient exchange waiting for rank 0's next launch,
    # and the loop has one D2H copy and one host synchronisation per iteration.  Evaluation runs on EVERY rank (same seed,
    # same parameters: same result) so that the ranks stay in step; rank 0 logs it.
    pending = None
    while trainer.step < args.total_timesteps:
        trainer.iteration()
        handle = trainer.stage_scalars()            # collective on multi-GPU runs: every rank calls it
        ev = None
        if (trainer.training_step / (args.epochs * args.num_minibatches)) % args.eval_steps == 0:   # MME:614
            ev = evaluate_fn(trainer, args.num_eval_ep, seed=args.seed + 7919 * trainer.training_step)
        if pending is not None:
            log(pending)
        pending = (handle, trainer.step, trainer.training_step, trainer.num_episodes, ev)
    if pending is not None:
        log(pending)
    if writer is not None:
        writer.close()
        if args.use_wnb:
            import wandb
            wandb.finish()
    return trainer


if __name__ == "__main__":
    main()
