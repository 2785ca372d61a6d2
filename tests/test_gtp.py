"""GTP front end (reference deep_mcts/gtp_interface.py): move parsing/formatting on the CPU, the
command loop and a GTPAgent-vs-engine game on the GPU."""
import numpy as np
import pytest


def test_move_parsing_and_formatting_round_trip():
    from deep_mcts_b200.gtp_interface import GTPInterface
    from deep_mcts_b200.hex.gtp_interface import HexGTPInterface
    from deep_mcts_b200.hex_with_swap.gtp_interface import HexWithSwapGTPInterface
    from deep_mcts_b200.othello.gtp_interface import OthelloGTPInterface
    for cls, g, extra in ((OthelloGTPInterface, 8, "pass"), (HexGTPInterface, 7, None), (HexWithSwapGTPInterface, 7, "swap")):
        for a in range(g * g):
            assert cls.parse_move(cls.format_move(a, g), g) == a
        assert cls.format_move(0, g) == "a1" and cls.format_move(g * g - 1, g) == f"{chr(ord('a') + g - 1)}{g}"
        assert cls.parse_move("C2", g) == 2 + 1 * g
        if extra:
            assert cls.parse_move(extra.upper(), g) == g * g and cls.format_move(g * g, g) == extra
        for bad in ("", "11", "z1", "a0", f"a{g + 1}"):
            with pytest.raises(ValueError):
                cls.parse_move(bad, g)
    assert GTPInterface.parse_player("B") == 1 and GTPInterface.parse_player("white") == 0
    assert GTPInterface.format_player(1) == "black" and GTPInterface.format_player(0) == "white"
    with pytest.raises(ValueError):
        GTPInterface.parse_player("green")


@pytest.mark.gpu
def test_gtp_command_loop_plays_a_game():
    from deep_mcts_b200.othello.gtp_interface import OthelloGTPInterface
    gtp = OthelloGTPInterface(4)
    gtp.num_simulations = 16
    assert gtp.run_command("name") == "Deep MCTS" and gtp.run_command("protocol_version") == "2"
    assert gtp.run_command("known_command genmove") == "true" and gtp.run_command("known_command foo") == "false"
    assert "genmove" in gtp.run_command("list_commands")
    with pytest.raises(ValueError, match="invalid command"):
        gtp.run_command("frobnicate")
    with pytest.raises(ValueError, match="play takes 2 arguments"):
        gtp.run_command("play b")
    assert gtp.run_command("boardsize 4 4") is None and gtp.run_command("clear_board") is None
    manager = gtp.game_manager
    state = manager.initial_game_state()
    legal = manager.legal_actions(state)
    illegal = next(a for a in range(16) if a not in legal)
    with pytest.raises(ValueError, match="illegal move"):
        gtp.run_command(f"play black {gtp.format_move(illegal, 4)}")
    # controller plays black's first legal move, the engine answers for white, and so on to the end
    plies = 0
    while not manager.is_final_state(gtp.mcts.state) and plies < 40:
        mover = gtp.format_player(gtp.mcts.state.player)
        if plies % 2 == 0:
            action = manager.legal_actions(gtp.mcts.state)[0]
            before = gtp.mcts.state
            assert gtp.run_command(f"play {mover} {gtp.format_move(action, 4)}") is None
            assert gtp.mcts.state == manager.generate_child_state(before, action)
        else:
            before = gtp.mcts.state
            move = gtp.run_command(f"genmove {mover}")
            action = gtp.parse_move(move, 4)
            assert action in manager.legal_actions(before)
            assert gtp.mcts.state == manager.generate_child_state(before, action)
        plies += 1
    assert manager.is_final_state(gtp.mcts.state)
    assert gtp.run_command("result") in ("Outcome.FIRST_PLAYER_WIN", "Outcome.SECOND_PLAYER_WIN", "Outcome.DRAW")
    assert "1" in gtp.run_command("showboard")
    # asking the "wrong" colour to move rewrites the mover, as the reference does
    gtp.run_command("clear_board")
    move = gtp.run_command("genmove white")
    assert gtp.parse_move(move, 4) in range(17)


@pytest.mark.gpu
def test_gtp_agent_plays_through_an_engine_process():
    from deep_mcts_b200.hex.game import HexManager
    from deep_mcts_b200.hex.gtp_interface import HexGTPInterface
    from deep_mcts_b200.gtp_interface import GTPAgent
    manager = HexManager(5)
    agent = GTPAgent(manager, 5, "deep_mcts_b200.hex.gtp_interface", HexGTPInterface.format_move,
                     HexGTPInterface.parse_move)
    try:
        state = manager.initial_game_state()
        rs = np.random.RandomState(0)
        for _ in range(3):
            action = agent.play(state)
            assert action in manager.legal_actions(state)
            state = manager.generate_child_state(state, action)
            if manager.is_final_state(state):
                break
            legal = manager.legal_actions(state)
            state = manager.generate_child_state(state, legal[rs.randint(len(legal))])
        agent.reset()
        assert agent.play(manager.initial_game_state()) in range(25)
    finally:
        agent.close()
