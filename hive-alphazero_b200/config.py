"""Constants of the Hive hot path, same names and values as the reference's
hive_engine/config.py:8-35 and woker/solo_play.py:23-30 (they are part of the contract)."""
import string

MAX_MAP_HAFT = 6
MAX_MAP_FULL = MAX_MAP_HAFT * 2                       # config.py:9
ACTION_SPACE = MAX_MAP_FULL * MAX_MAP_FULL * 11       # config.py:10 -> 1584
STATE_FEATURES = 56                                   # config.py:21
MAX_GAME_LENGTH = 55                                  # config.py:23
MAX_LEN_BACK = 5
SEARCH_THREADS = 32
MAX_PROCESS = 60
BOT_WEIGHT = 0.24
LOSS_WEIGHT = {"value": 1.0, "policy": 1.0}
DISCOUNTED_REWARD = 0.99

# cell labels (config.py:12-17): index_char H..S, index_number 7..18
index_number = [str(i) for i in range(1, 27)][12 - MAX_MAP_HAFT:12 + MAX_MAP_HAFT]
index_char = list(string.ascii_uppercase)[13 - MAX_MAP_HAFT:13 + MAX_MAP_HAFT]

# settings.py:3-4
PIECE_WHITE = (250, 250, 250)
PIECE_BLACK = (71, 71, 71)
WIDTH = 1400
HEIGHT = 900 + 250

# woker/solo_play.py:23-30
simulation_num_per_move = 100
tau_decay_rate = 0.01
c_puct = 0.7
dirichlet_alpha = 0.3
noise_eps = 0.25
virtual_loss = 1

# piece order per colour (inventory_frame.py:47-99); keys as in env_hive.py:71-87 are
# str(type)+index -- we expose the short ids the reference uses in state_key (env_hive.py:79,86)
PIECE_IDS = ["Q0", "B0", "B1", "S0", "S1", "G0", "G1", "G2", "A0", "A1", "A2"]
PIECE_CLASS = ["Queen", "Beetle", "Beetle", "Spider", "Spider", "Grasshopper", "Grasshopper", "Grasshopper",
               "Ant", "Ant", "Ant"]
PIECE_NUM = [0, 0, 1, 0, 1, 0, 1, 2, 0, 1, 2]
HAND = 255
NOOP = -2
