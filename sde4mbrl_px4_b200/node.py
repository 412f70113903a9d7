"""ROS-free harness that drives the controller interface exactly the way the reference
node does.

It restates the *calling pattern* of ``SDEControlROS``
(/root/reference sde4mbrl_px4/mpc_controller/sde_control.py): two controllers (trajectory
tracker + set-point controller, :156-177), a solver process forked after the controllers
are built (:723-728) and fed through shared memory + an event (:616-663), the control
automaton none/pos/idle/traj (:180-220), the per-state-message callback that picks
``u_opt[index]`` from the last finished plan (:223-325) and the two service callbacks
(:453-562).  MAVLink / ROS transport is replaced by plain Python calls: a state message
is a dict with the ``MPC_FULL_STATE`` fields, a command is a dict with the
``MPC_MOTORS_CMD`` fields (:605-613).
"""
from __future__ import annotations

import multiprocessing
import os
import time
from multiprocessing import shared_memory

import numpy as np

from . import sde_mpc_design as design
from .utils import enu2ned

CONTROL_STATES = {"none": 0, "reset": 1, "test": 2, "pos": 3, "idle": 4, "traj": 5}   # sde_control.py:46
STATE_NAMES = {v: k for k, v in CONTROL_STATES.items()}
# srv/FollowTraj.srv controller_state enum
CTRL_INACTIVE, CTRL_TRAJ_ACTIVE, CTRL_TRAJ_IDLE, CTRL_POSE_ACTIVE, CTRL_TEST = 0, 1, 2, 3, 4
INFO_KEYS = ("sample_time_posmpc", "avg_linesearch", "stepsize", "num_steps", "grad_norm", "avg_stepsize",
             "cost0", "costT", "solveTime")   # sde_control.py:647-648


def dummy_state() -> np.ndarray:
    """sde_control.py:746-747"""
    return np.array([0, 0, 0, 0, 0, 0, 1, 0, 0, 0, 0, 0, 0], dtype=np.float32)


class _Shm:
    """A named float array in POSIX shared memory (sde_control.py:622-655)."""

    def __init__(self, shape, dtype):
        n = int(np.prod(shape)) * np.dtype(dtype).itemsize
        self.mem = shared_memory.SharedMemory(create=True, size=max(n, 8))
        self.arr = np.ndarray(shape, dtype=dtype, buffer=self.mem.buf)
        self.arr[...] = 0

    def close(self):
        self.arr = None
        try:
            self.mem.close()
            self.mem.unlink()
        except Exception:
            pass


class SDEControlNode:
    def __init__(self, config_dir: str, traj_ctrl: str, sp_ctrl: str, seed: int = 10, mpc_report_dt: float = 0.2,
                 use_process: bool = True, loader=None, clock=time.time, device: int = 0, **overrides):
        self.seed, self.mpc_report_dt, self.clock = int(seed), mpc_report_dt, clock
        self._loader = loader or (lambda p, convert_to_enu=True: design.load_mpc_from_cfgfile(
            p, convert_to_enu=convert_to_enu, device=device, **overrides))
        self._control_state = CONTROL_STATES["none"]
        self._target_x = dummy_state()
        self._run_trajectory, self._trajec_time, self._pos_control, self._test_mode = False, -1.0, False, False
        self.mpc_on = CONTROL_STATES["none"]
        self.last_traj_time = 0.0
        self._index = 0
        self.current_weight_motors = 0
        self.reset_done = False
        self._target_sp = np.array([0, 0, 0, 1, 0, 0, 0], np.float32)
        self.last_time_state_info = None
        self.dt_state_callback = self.dt_state_info = 0.0
        self.last_command = None
        self.load_times = {}
        self._load_mpc_models(os.path.join(config_dir, traj_ctrl), os.path.join(config_dir, sp_ctrl))
        self._shared_variables()
        self.use_process = use_process
        self._proc = None
        if use_process:
            self._start_process()
        else:
            self._loop_state = None

    # ------------------------------------------------------------------ loading (:156-177, :681-721)
    def _load_single(self, path):
        x0 = dummy_state()
        cfg_dict, (m_reset, m_mpc), state_from_traj, _ = self._loader(path, convert_to_enu=True)
        t_init = 0.01
        sft = None
        if state_from_traj is not None:
            sft = design.jit(state_from_traj).lower(t_init).compile()
        rng = design.PRNGKey(self.seed)
        t = time.time()
        reset = design.jit(m_reset).lower(x=x0, rng=rng, xdes=x0).compile()
        opt_state0 = reset(x=x0, rng=rng, xdes=x0)
        opt_state0.yk.block_until_ready()
        self.load_times[path] = time.time() - t
        mpc = design.jit(m_mpc).lower(x0, rng, opt_state0, curr_t=t_init, xdes=x0).compile()
        # The reference also evaluates the solver once here, in the parent.  With a CUDA
        # back end that would create a context before the fork, so the
        # first evaluation is left to the solver process; the default plan is the reset plan.
        uopt0 = np.array(opt_state0.yk)
        return sft, reset, mpc, uopt0, opt_state0, cfg_dict

    def _load_mpc_models(self, traj_path, pos_path):
        (self.state_from_traj, self.reset_traj_mpc, self.mpc_traj_solver, self.traj_uopt, self.default_traj_opt_state,
         self.traj_cfg_dict) = self._load_single(traj_path)
        assert self.state_from_traj is not None, "the trajectory controller config must provide trajectory_path"
        self._dt_usec_traj = self.traj_cfg_dict["_time_steps"][0] * 1e6
        (sft_pos, self.reset_pos_mpc, self.mpc_pos_solver, self.pos_uopt, self.default_pos_opt_state,
         self.pos_cfg_dict) = self._load_single(pos_path)
        assert sft_pos is None, "the set-point controller config must not provide trajectory_path"
        self._dt_usec_pos = self.pos_cfg_dict["_time_steps"][0] * 1e6
        self._dt_usec = self._dt_usec_pos

    # ------------------------------------------------------------------ shared variables (:616-663)
    def _shared_variables(self):
        u_zero = self.traj_uopt if self.traj_uopt.nbytes > self.pos_uopt.nbytes else self.pos_uopt
        self._shm = dict(
            curr_state=_Shm((13,), np.float32), u_opt=_Shm(u_zero.shape, np.float32),
            w_opt=_Shm((u_zero.shape[0], 4), np.float64), info_pre=_Shm((3,), np.float64),
            opt_info=_Shm((len(INFO_KEYS),), np.float32), target=_Shm((13,), np.float32))
        s = self.default_traj_opt_state
        self._shm["opt_info"].arr[:] = [-1.0, s.avg_linesearch, s.stepsize, s.num_steps, s.grad_sqr, s.avg_stepsize,
                                        s.init_cost, s.opt_cost, 0.0]
        self._shm["info_pre"].arr[:] = [0.0, 0.0, CONTROL_STATES["none"]]
        self._optimizer_info = np.array(self._shm["opt_info"].arr)
        ctx = multiprocessing.get_context("fork")
        self._curr_state_lock, self._u_opt_lock = ctx.Lock(), ctx.Lock()
        self._mpc_event, self._done_event = ctx.Event(), ctx.Event()
        self._ctx = ctx

    def _start_process(self):
        self._proc = self._ctx.Process(target=self._mpc_process_fn, name="mpc_process", daemon=True)
        self._proc.start()

    def close(self):
        if self._proc is not None:
            self._proc.terminate()
            self._proc.join(timeout=5)
            self._proc = None
        for s in self._shm.values():
            s.close()

    # ------------------------------------------------------------------ solver loop (:328-450)
    def _loop_init(self):
        rng = design.PRNGKey(self.seed)
        _, rng_traj, rng_pos = design.split(rng, 3)
        x0 = dummy_state()
        st = dict(rng_traj=rng_traj, rng_pos=rng_pos, curr_ctrl=None, idle_traj=False,
                  opt_traj=self.reset_traj_mpc(x=x0, rng=rng_traj, xdes=x0),
                  opt_pos=self.reset_pos_mpc(x=x0, rng=rng_pos, xdes=x0))
        # one warm-up evaluation of each solver (:349-350); creates the CUDA context HERE,
        # i.e. in the solver process
        self.mpc_traj_solver(x0, rng_traj, st["opt_traj"], curr_t=0.0, xdes=x0)
        self.mpc_pos_solver(x0, rng_pos, st["opt_pos"], curr_t=0.0, xdes=x0)
        return st

    def _loop_once(self, st):
        sh = self._shm
        with self._curr_state_lock:
            curr_state = np.array(sh["curr_state"].arr)
            sample_time, trajec_time, control_state = sh["info_pre"].arr[0], sh["info_pre"].arr[1], int(sh["info_pre"].arr[2])
            target_x = np.array(sh["target"].arr) if control_state != CONTROL_STATES["traj"] else None
        t0 = time.time()
        cs = CONTROL_STATES
        if st["curr_ctrl"] is None or (st["curr_ctrl"] == "none" and control_state != cs["none"]):
            st["opt_traj"] = self.reset_traj_mpc(x=curr_state, rng=st["rng_traj"], xdes=curr_state)
            st["opt_pos"] = self.reset_pos_mpc(x=curr_state, rng=st["rng_pos"], xdes=curr_state)
        if control_state == cs["idle"] and st["curr_ctrl"] in (None, "none", "pos"):
            st["opt_traj"] = self.reset_traj_mpc(x=curr_state, rng=st["rng_traj"], xdes=curr_state)
            st["curr_ctrl"], st["idle_traj"] = "idle", True
        if control_state == cs["none"]:
            st["curr_ctrl"] = "none"
            uopt, st["opt_pos"], st["rng_pos"], x_evol = self.mpc_pos_solver(
                curr_state, st["rng_pos"], st["opt_pos"], curr_t=0.0, xdes=enu2ned(curr_state, np))
        elif control_state == cs["idle"]:
            st["curr_ctrl"] = "idle"
            uopt, st["opt_pos"], st["rng_pos"], x_evol = self.mpc_pos_solver(
                curr_state, st["rng_pos"], st["opt_pos"], curr_t=0.0, xdes=target_x)
            st["idle_traj"] = not st["idle_traj"]
            if st["idle_traj"]:   # every second tick also warm the trajectory solver (:406-408)
                _, st["opt_traj"], st["rng_traj"], _ = self.mpc_traj_solver(
                    curr_state, st["rng_traj"], st["opt_traj"], curr_t=trajec_time, xdes=curr_state)
        elif control_state == cs["traj"]:
            st["curr_ctrl"] = "traj"
            uopt, st["opt_traj"], st["rng_traj"], x_evol = self.mpc_traj_solver(
                curr_state, st["rng_traj"], st["opt_traj"], curr_t=trajec_time, xdes=curr_state)
        elif control_state == cs["pos"]:
            st["curr_ctrl"] = "pos"
            uopt, st["opt_pos"], st["rng_pos"], x_evol = self.mpc_pos_solver(
                curr_state, st["rng_pos"], st["opt_pos"], curr_t=0.0, xdes=target_x)
        else:
            raise ValueError(f"Unknown control state: {control_state}")
        uopt.block_until_ready()
        solve_time = time.time() - t0
        uopt = np.array(uopt)
        thrust = np.sum(uopt, axis=1) / uopt.shape[1]                                   # :431
        wopt = np.array([thrust, x_evol[1:, 10], x_evol[1:, 11], x_evol[1:, 12]]).T      # :432
        o = st["opt_traj"] if st["curr_ctrl"] in ("traj", "idle") else st["opt_pos"]
        with self._u_opt_lock:
            sh["u_opt"].arr[: uopt.shape[0], : uopt.shape[1]] = uopt
            sh["w_opt"].arr[: wopt.shape[0], :] = wopt
            sh["opt_info"].arr[:] = [sample_time, float(o.avg_linesearch), float(o.stepsize), float(o.num_steps),
                                     float(o.grad_sqr), float(o.avg_stepsize), float(o.init_cost), float(o.opt_cost),
                                     solve_time]
        return solve_time

    def _mpc_process_fn(self):
        st = self._loop_init()
        while True:
            self._mpc_event.wait()
            self._mpc_event.clear()
            self._loop_once(st)
            self._done_event.set()

    def wait_solve(self, timeout: float = 30.0) -> bool:
        """Block until the solver process has finished the solve requested by the last state
        message (test/benchmark helper; the flight path never waits)."""
        if not self.use_process:
            return True
        ok = self._done_event.wait(timeout)
        self._done_event.clear()
        return ok

    # ------------------------------------------------------------------ automaton (:180-220)
    def control_automata(self) -> int:
        cs = CONTROL_STATES
        if self._pos_control:
            p = self._target_sp        # (x, y, z, qw, qx, qy, qz), ENU (:187-192)
            self._target_x = np.array([p[0], p[1], p[2], 0, 0, 0, p[3], p[4], p[5], p[6], 0, 0, 0], dtype=np.float32)
            self._dt_usec = self._dt_usec_pos
            return cs["pos"]
        if self._trajec_time < 0.0:
            self._dt_usec = self._dt_usec_pos
            return cs["none"]
        if not self._run_trajectory:
            self._trajec_time = 0
            self._target_x = np.array(self.state_from_traj(0.0), dtype=np.float32)
            self._dt_usec = self._dt_usec_pos
            return cs["idle"]
        now = self.clock()
        self._dt_usec = self._dt_usec_traj
        if self._trajec_time == 0:
            self.last_traj_time = now
            self._trajec_time = 0.0000001
        else:
            self._trajec_time = now - self.last_traj_time
        return cs["traj"]

    # ------------------------------------------------------------------ state callback (:223-325)
    def mpc_state_callback(self, msg: dict):
        """``msg``: MPC_FULL_STATE fields time_usec,x,y,z,vx,vy,vz,qw,qx,qy,qz,wx,wy,wz.
        Returns the MPC_MOTORS_CMD dict that would be sent, or None."""
        t_in = time.time()
        if self.last_time_state_info is not None:
            self.dt_state_info = t_in - self.last_time_state_info
        self.last_time_state_info = t_in
        curr_state = np.array([msg[k] for k in ("x", "y", "z", "vx", "vy", "vz", "qw", "qx", "qy", "qz", "wx", "wy", "wz")],
                              dtype=np.float32)
        self.sample_time = msg["time_usec"]
        self._control_state = self.control_automata()
        sh = self._shm
        with self._curr_state_lock:
            sh["curr_state"].arr[:] = curr_state
            sh["info_pre"].arr[:] = [msg["time_usec"], float(self._trajec_time), self._control_state]
            if self._control_state != CONTROL_STATES["traj"]:
                sh["target"].arr[:] = self._target_x
        if self.use_process:
            self._mpc_event.set()
        else:
            if self._loop_state is None:
                self._loop_state = self._loop_init()
            self._loop_once(self._loop_state)
        with self._u_opt_lock:
            u_opt, w_opt = np.array(sh["u_opt"].arr), np.array(sh["w_opt"].arr)
            self._optimizer_info = np.array(sh["opt_info"].arr)
        tsample_mpc = self._optimizer_info[0]
        if tsample_mpc <= 0:          # no MPC solution computed yet (:284-288)
            self.dt_state_callback = time.time() - t_in
            return None
        index = int((self.sample_time - tsample_mpc) / self._dt_usec)
        u_rows = self.traj_uopt.shape[0] if self._control_state == CONTROL_STATES["traj"] else self.pos_uopt.shape[0]
        if index >= u_rows:           # plan exhausted: keep the last control (:294-298)
            index = u_rows - 1
        index = max(index, 0)
        uopt = u_opt[index, :]
        if uopt.shape[0] < 6:         # zero-pad to 6 motors (:302-303)
            uopt = np.concatenate((uopt, np.zeros((6 - uopt.shape[0],), uopt.dtype)))
        self._uopt, self._wopt, self._index = uopt, w_opt[index, :], index
        if self._control_state == CONTROL_STATES["none"]:
            self.dt_state_callback = time.time() - t_in
            return None
        self.mpc_on = self._control_state if not self._test_mode else CONTROL_STATES["test"]
        cmd = dict(time_usec=int(self.clock() * 1e6), motor_val_des=self._uopt.copy(),
                   thrust_and_angrate_des=self._wopt.copy(), mpc_on=self.mpc_on, weight_motors=self.current_weight_motors)
        self.last_command = cmd
        self.dt_state_callback = time.time() - t_in
        return cmd

    # ------------------------------------------------------------------ MAVLink edge (:142-154, 605-613)
    def feed_mavlink(self, data: bytes) -> bytes:
        """Byte-level edge of the node: ``data`` is what mavlink-router delivers (scripts/router_sitl.conf:18-19).  Every
        complete MPC_FULL_STATE frame (id 367) goes through ``mpc_state_callback`` — the body of the reference's
        ``read_mavlink_msg`` loop (:142-154) — and every resulting command is returned as an MPC_MOTORS_CMD frame (id 368),
        what ``pub_cmd_setpoint`` (:605-613) sends.  Other message ids and corrupted frames are skipped."""
        from . import mavlink_codec as mv

        if getattr(self, "_mav_decoder", None) is None:
            self._mav_decoder, self._mav_seq = mv.Decoder(), 0
        out = b""
        for msg, values, _ in self._mav_decoder.feed(data):
            if msg.msgid != mv.MPC_FULL_STATE.msgid:
                continue
            cmd = self.mpc_state_callback(values)
            if cmd is not None:
                out += mv.encode(mv.MPC_MOTORS_CMD, mv.motors_cmd_values(cmd["time_usec"], cmd["motor_val_des"], cmd["thrust_and_angrate_des"],
                                                                        cmd["mpc_on"], cmd["weight_motors"]), seq=self._mav_seq)
                self._mav_seq = (self._mav_seq + 1) & 0xFF
        return out

    # ------------------------------------------------------------------ services (:453-562)
    def initialize_mpc(self) -> bool:
        """``controller_init`` -> set_trajectory_and_params (:453-477): refused while a controller runs;
        otherwise marks the controller as reset (the node also sends mpc_on='reset' five times)."""
        if self._run_trajectory or self._pos_control:
            return False
        self.mpc_on = CONTROL_STATES["reset"]
        self.reset_done = True
        return True

    def start_trajectory(self, state_controller: int, target_pose=None, weight_motors: int = 255) -> bool:
        """``start_trajectory`` service (srv/FollowTraj.srv; :480-562).
        ``target_pose`` = (x, y, z, qw, qx, qy, qz) in ENU, used by the set-point modes."""
        if 0 <= weight_motors <= 100:      # weight-only update (:485-488)
            self.current_weight_motors = int(weight_motors)
            return True
        if not getattr(self, "reset_done", False) and state_controller != CTRL_INACTIVE:
            return False                   # controller_init first (:491-494)
        if target_pose is not None:
            self._target_sp = np.asarray(target_pose, np.float32)
        cs = CONTROL_STATES
        if state_controller == CTRL_TEST:          # controller_test (:500-510)
            self.mpc_on = cs["test"]
            self._test_mode, self._pos_control, self._run_trajectory, self._trajec_time = True, True, False, -1.0
            return True
        if state_controller == CTRL_POSE_ACTIVE:   # (:513-522)
            self.mpc_on = cs["pos"]
            self._test_mode, self._pos_control, self._run_trajectory, self._trajec_time = False, True, False, -1.0
            return True
        if state_controller == CTRL_INACTIVE:      # (:524-539)
            self.reset_done = False
            self.mpc_on = cs["none"]
            self._test_mode, self._pos_control, self._run_trajectory, self._trajec_time = False, False, False, -1.0
            return True
        if self._run_trajectory and state_controller == CTRL_TRAJ_ACTIVE:
            return False                   # already running (:542-545)
        self._trajec_time = 0.0 if state_controller in (CTRL_TRAJ_IDLE, CTRL_TRAJ_ACTIVE) else -1.0
        # TRAJ_ACTIVE is only honoured from idle; otherwise the controller goes to idle first (:548-553)
        self._run_trajectory = (self._control_state == cs["idle"] and state_controller == CTRL_TRAJ_ACTIVE)
        self._test_mode = self._pos_control = False
        self.mpc_on = cs["traj"] if self._run_trajectory else cs["idle"]
        return True

    # ------------------------------------------------------------------ telemetry (:564-585; msg/OptMPCState.msg)
    def opt_state_report(self) -> dict:
        i = dict(zip(INFO_KEYS, [float(v) for v in self._optimizer_info]))
        return dict(avg_linesearch=i["avg_linesearch"], avg_stepsize=i["avg_stepsize"], stepsize=i["stepsize"],
                    grad_norm=i["grad_norm"], cost_init=i["cost0"], opt_cost=i["costT"], num_steps=int(i["num_steps"]),
                    solve_time=i["solveTime"], callback_dt=self.dt_state_callback, state_dt=self.dt_state_info,
                    ctrl_state=STATE_NAMES[self._control_state], mpc_indx=int(self._index))
