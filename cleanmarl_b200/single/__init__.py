"""Drop-ins for the reference's single-env PPO scripts (mappo.py, ippo.py, mappo_lstm.py, ippo_lstm.py)."""
