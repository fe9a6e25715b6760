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
    while trainer.step < args.total_timesteps:
        trainer.iteration()
        step = trainer.step
        roll = trainer.rollout_scalars()            # collective: every rank calls it
        pending_episodes += args.batch_size
        pending_reward_sum += roll["ep_reward"] * args.batch_size
        if writer is not None:
            if pending_episodes > args.log_every:
                writer.add_scalar("rollout/ep_reward", pending_reward_sum / pending_episodes, step)
                writer.add_scalar("rollout/ep_length", roll["ep_length"], step)
                writer.add_scalar("rollout/num_episodes", trainer.num_episodes, step)
                pending_episodes, pending_reward_sum = 0, 0.0
            for k, v in trainer.train_scalars().items():            # MME:605-612
                writer.add_scalar(f"train/{k}", v, step)
            writer.add_scalar("train/num_updates", trainer.training_step, step)
            if (trainer.training_step / args.epochs) % args.eval_steps == 0:       # MME:614
                mean, std, length = evaluate_fn(trainer, args.num_eval_ep, seed=args.seed + 7919 * trainer.training_step)
                writer.add_scalar("eval/ep_reward", mean, step)
                writer.add_scalar("eval/std_ep_reward", std, step)
                writer.add_scalar("eval/ep_length", length, step)
    if writer is not None:
        writer.close()
        if args.use_wnb:
            import wandb
            wandb.finish()
    return trainer


if __name__ == "__main__":
    main()
