"""CPU oracle for the rollout-and-update hot path — TEST INFRASTRUCTURE ONLY.

Nothing under ``gymnasium_solver_b200/`` may import this package.  Only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl reference`` legs use it,
and only as the checker or as the timed CPU baseline — never as the product path.

Modules
    envs     ctypes front-end of ``envs.c`` (fp64 restatement of gymnasium 1.1.1 classic control +
             TimeLimit + SyncVectorEnv NEXT_STEP autoreset + RecordEpisodeStatistics + reference wrappers)
    returns  numpy restatement of the reference's utils/returns_advantages.py
    policy   torch fp32 restatement of utils/models.py MLP forward, Categorical / MaskedCategorical and
             the PPO / REINFORCE losses of agents/ppo/ppo_agent.py, agents/reinforce/reinforce_agent.py
"""
