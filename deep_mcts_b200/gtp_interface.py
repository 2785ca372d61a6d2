"""Go Text Protocol front end on the batch-width-1 MCTS (interface of reference
deep_mcts/gtp_interface.py:20-238: ``GTPInterface`` with the same command table, argument
checks and error strings, and ``GTPAgent`` that plays through a GTP engine process).

What changes underneath: the tree lives on the device, so ``play`` re-roots with
``MCTS.observe`` (child of the current root when the move was searched, a fresh root
otherwise -- the reference's ``root.children.get(action, Node())``, gtp_interface.py:114-116),
and ``GTPAgent`` talks to the engine over plain pipes (``subprocess``) instead of pexpect.
"""
import dataclasses
import os
import re
import subprocess
import sys
from abc import ABC, abstractmethod
from typing import Callable, Dict, List, NoReturn, Optional

import numpy as np

from .game import Action, GameManager, Player
from .gamenet import GameNet, cached_state_evaluator
from .mcts import MCTS
from .tournament import Agent


class GTPInterface(ABC):
    commands: Dict[str, Callable[[List[str]], Optional[str]]]
    num_simulations = 100

    def __init__(self, board_size: int) -> None:
        self.commands = {
            "name": self.name,
            "version": self.version,
            "protocol_version": self.protocol_version,
            "known_command": self.known_command,
            "list_commands": self.list_commands,
            "quit": self.quit,
            "boardsize": self.boardsize,
            "clear_board": self.clear_board,
            "play": self.play,
            "genmove": self.genmove,
            "showboard": self.showboard,
            "result": self.result,
        }
        self.board_size = board_size
        self.net: Optional[GameNet] = None
        self.clear_board([])

    # ------------------------------------------------------------------ dispatch
    def run_command(self, command: str) -> Optional[str]:
        command, *args = command.split()
        if command not in self.commands:
            raise ValueError("invalid command")
        return self.commands[command](args)

    def start(self) -> None:
        """Read commands from stdin, answer ``= result`` / ``? error`` (gtp_interface.py:183-196)."""
        while True:
            command = input("")
            if not command:
                continue
            try:
                result = self.run_command(command)
            except ValueError as e:
                print(f"? {e}\n", flush=True)
            else:
                print("= \n" if result is None else f"= {result}\n", flush=True)

    # ------------------------------------------------------------ administrative
    def name(self, args: List[str]) -> str:
        return "Deep MCTS"

    def version(self, args: List[str]) -> str:
        return "0.0.1"

    def protocol_version(self, args: List[str]) -> str:
        return "2"

    def known_command(self, args: List[str]) -> str:
        if len(args) != 1:
            raise ValueError("known_command takes 1 argument")
        return str(args[0] in self.commands).lower()

    def list_commands(self, args: List[str]) -> str:
        return "\n" + "\n".join(self.commands)

    def quit(self, args: List[str]) -> NoReturn:
        sys.exit(0)

    # ------------------------------------------------------------------- set-up
    def boardsize(self, args: List[str]) -> None:
        if len(args) not in (1, 2):   # HexGui sends the size twice
            raise ValueError("boardsize takes 1 argument")
        try:
            board_size = int(args[0])
        except ValueError:
            raise ValueError("invalid board size")
        if board_size < 1:
            raise ValueError("invalid board size")
        self.board_size = board_size
        self.clear_board([])

    def clear_board(self, args: List[str]) -> None:
        if self.net is None or self.net.manager.grid_size != self.board_size:
            self.net = self.get_game_net(self.board_size)
        self.game_manager: GameManager = self.net.manager
        self.mcts = MCTS(self.game_manager, self.num_simulations, None, cached_state_evaluator(self.net))

    # --------------------------------------------------------------------- play
    def _set_mover(self, player: int) -> None:
        """The controller may ask either colour to move: rewrite the mover (and drop the tree)."""
        if int(self.mcts.state.player) != int(player):
            self.mcts.observe(dataclasses.replace(self.mcts.state, player=Player(int(player))))

    def play(self, args: List[str]) -> None:
        if len(args) != 2:
            raise ValueError("play takes 2 arguments")
        player = self.parse_player(args[0])
        action = self.parse_move(args[1], self.board_size)
        self._set_mover(player)
        if action not in self.game_manager.legal_actions(self.mcts.state):
            raise ValueError("illegal move")
        self.mcts.observe(self.game_manager.generate_child_state(self.mcts.state, action))

    def genmove(self, args: List[str]) -> str:
        if len(args) != 1:
            raise ValueError("play takes 1 argument")
        self._set_mover(self.parse_player(args[0]))
        action_probabilities, _ = self.mcts.step()
        value, net_action_probabilities = self.net.evaluate(self.mcts.state)
        print(value, file=sys.stderr)
        print(self.game_manager.probabilities_grid(net_action_probabilities.tolist()), file=sys.stderr)
        print(file=sys.stderr)
        print(self.game_manager.probabilities_grid(action_probabilities), file=sys.stderr)
        action = int(np.argmax(action_probabilities))
        self.mcts.advance(action)
        return self.format_move(action, self.board_size)

    def showboard(self, args: List[str]) -> str:
        return f"\n{self.mcts.state}"

    def result(self, args: List[str]) -> str:
        return str(self.game_manager.evaluate_final_state(self.mcts.state))

    # ------------------------------------------------------------------ parsing
    @staticmethod
    def parse_player(player: str) -> int:
        player = {"w": "white", "b": "black"}.get(player.lower(), player.lower())
        if player not in ("white", "black"):
            raise ValueError("invalid player")
        return Player.FIRST if player == "black" else Player.SECOND

    @staticmethod
    def format_player(player: int) -> str:
        return "black" if player == Player.FIRST else "white"

    @staticmethod
    @abstractmethod
    def parse_move(move: str, board_size: int) -> Action: ...

    @staticmethod
    @abstractmethod
    def format_move(move: Action, board_size: int) -> str: ...

    @staticmethod
    @abstractmethod
    def get_game_net(board_size: int) -> GameNet: ...


_CELL_RE = re.compile(r"([a-z])(\d{1,2})")


def parse_cell(move: str, board_size: int) -> Action:
    """``<column letter><row number>`` -> ``x + y * board_size`` (othello/gtp_interface.py:17-30)."""
    match = _CELL_RE.match(move.lower())
    if match is None:
        raise ValueError("invalid move")
    x, y = ord(match.group(1)) - ord("a"), int(match.group(2)) - 1
    if x >= board_size or not 0 <= y < board_size:
        raise ValueError("invalid move")
    return x + y * board_size


def format_cell(move: Action, board_size: int) -> str:
    return f"{chr(ord('a') + move % board_size)}{move // board_size + 1}"


def net_from_argv(net_cls, board_size: int):
    """``<interface> <board size> <checkpoint>``: load a .pth state_dict or a full checkpoint when the
    requested size matches, else a fresh network (othello/gtp_interface.py:38-46)."""
    if len(sys.argv) == 3 and board_size == int(sys.argv[1]):
        path = sys.argv[2]
        return net_cls.from_path(path, board_size) if path.endswith(".pth") else net_cls.from_path_full(path)
    return net_cls(board_size)


def board_gtp_interface(name: str, net_cls, default_size: int, special_move: Optional[str] = None,
                        hexgui: bool = False):
    """The per-game engines differ in three facts: the network class, the name of the extra action
    ``board_size ** 2`` ("pass" for Othello, "swap" for Hex with swap, none for Hex) and whether HexGui's
    ``hexgui-analyze_commands`` probe is answered.  Returns the ``GTPInterface`` subclass."""

    def parse_move(move: str, board_size: int) -> Action:
        if special_move is not None and move.lower() == special_move:
            return board_size ** 2
        return parse_cell(move, board_size)

    def format_move(move: Action, board_size: int) -> str:
        if special_move is not None and move == board_size ** 2:
            return special_move
        return format_cell(move, board_size)

    def __init__(self, board_size: int = default_size) -> None:
        GTPInterface.__init__(self, board_size)
        if hexgui:
            self.commands["hexgui-analyze_commands"] = lambda args: None

    body = {"__init__": __init__, "parse_move": staticmethod(parse_move), "format_move": staticmethod(format_move),
            "get_game_net": staticmethod(lambda board_size: net_from_argv(net_cls, board_size)),
            "default_size": default_size, "__doc__": f"GTP engine for {net_cls.__name__}."}
    return type(name, (GTPInterface,), body)


def run_engine(interface_cls) -> None:
    """``python -m deep_mcts_b200.<game>.gtp_interface [<board size> <checkpoint>]``."""
    size = int(sys.argv[1]) if len(sys.argv) == 3 else interface_cls.default_size
    interface_cls(board_size=size).start()


class GTPAgent(Agent):
    """Tournament agent that plays through a GTP engine in a child process
    (reference gtp_interface.py:199-238).  ``engine_module`` is the module run with ``python -m``."""

    def __init__(self, manager: GameManager, grid_size: int, engine_module: str,
                 format_move: Callable[[Action, int], str], parse_move: Callable[[str, int], Action]) -> None:
        root = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
        env = dict(os.environ, PYTHONPATH=root + os.pathsep + os.environ.get("PYTHONPATH", ""))
        self.process = subprocess.Popen([sys.executable, "-m", engine_module], stdin=subprocess.PIPE,
                                        stdout=subprocess.PIPE, env=env, text=True, bufsize=1)
        self.format_move, self.parse_move = format_move, parse_move
        self.game_manager = manager
        self.grid_size = grid_size
        self._command(f"boardsize {grid_size}")
        self.state = manager.initial_game_state()

    def _command(self, command: str) -> str:
        self.process.stdin.write(command + "\n")
        self.process.stdin.flush()
        lines: List[str] = []
        while True:
            line = self.process.stdout.readline()
            if line == "":
                raise RuntimeError("GTP engine exited")
            if line.strip() == "" and lines:
                break
            if line.strip():
                lines.append(line.rstrip("\n"))
        if lines[0].startswith("?"):
            raise ValueError(lines[0][1:].strip())
        return lines[0][1:].strip()

    def play(self, state) -> Action:
        if self.state != state:
            action = next(a for a in self.game_manager.legal_actions(self.state)
                          if self.game_manager.generate_child_state(self.state, a) == state)
            self._command(f"play {GTPInterface.format_player(self.state.player)} "
                          f"{self.format_move(action, self.grid_size)}")
            self.state = state
        move = self._command(f"genmove {GTPInterface.format_player(self.state.player)}")
        action = self.parse_move(move, self.grid_size)
        self.state = self.game_manager.generate_child_state(self.state, action)
        return action

    def reset(self) -> None:
        self._command("clear_board")
        self.state = self.game_manager.initial_game_state()

    def close(self) -> None:
        if self.process.poll() is None:
            try:
                self.process.stdin.write("quit\n")
                self.process.stdin.flush()
                self.process.wait(timeout=10)
            except Exception:
                self.process.kill()
