"""Mirror of the reference's config.py (names kept 1:1, /root/reference/config.py:1-54) as one
frozen dataclass, convertible to the C ABI's ngp_config."""
from __future__ import annotations

from dataclasses import dataclass, field, replace
from typing import Tuple

from . import _lib


@dataclass(frozen=True)
class Config:
    BG_COLOUR: Tuple[int, int, int] = (144, 72, 17)
    BALL_COLOUR: Tuple[int, int, int] = (236, 236, 236)
    LEFT_GUY_COLOUR: Tuple[int, int, int] = (213, 130, 74)
    RIGHT_GUY_COLOUR: Tuple[int, int, int] = (92, 186, 92)
    GAME_BOTTOM: int = 194
    GAME_TOP: int = 34
    SCALED_PADDLE_HEIGHT: float = 16.0
    FPS: int = 60
    GAME_WIDTH: int = 160
    RIGHT_ACTION_START: int = 4
    RIGHT_ACTION_END: int = 6
    LEFT_ACTION_END: int = 8
    LEFT_PLAYER_START_BUTTON: int = -1
    RIGHT_PLAYER_START_BUTTON: int = 0
    N_CLASSES: int = 2
    TIMEOUT_THRESH: int = 2_000
    NETWORK_SHAPE: Tuple[int, ...] = (6, 2, 2)
    BIAS: bool = True
    PROBABILITY_OF_MUTATING_A_SINGLE_GENE: float = 0.9
    GAUSSIAN_MUTATION_SIGMA: float = 0.9
    GAUSSIAN_MUTATION_MEAN: float = 0.0
    GAUSSIAN_MUTATION_PROBABILITY: float = 0.9
    CROSSOVER_BLEND_PROBABILITY: float = 0.9
    CROSSOVER_BLEND_ALPHA: float = 0.9
    GENERATIONS_BEFORE_SAVE: int = 5
    GAMES_TO_PLAY: int = 6
    POPULATION_SIZE: int = 64
    WIN_SCORE: int = 3
    TIME_SCALER: float = 100.0
    # not in the reference: opponent schedule of one evaluation and a per-episode frame cap
    SCHEDULE: int = _lib.SCHEDULE_REFERENCE
    MAX_FRAMES: int = 0
    CORE: int = _lib.CORE_TRANSLATED           # 6507 core of the fused rollout
    # what each gym-retro button does to the console (include/ngp.h NGP_BTN_*); default = the reference's own names:
    # [0] RIGHT_PLAYER_START_BUTTON -> fire of paddle 1, [15] LEFT_PLAYER_START_BUTTON -> fire of paddle 0,
    # [4]/[5] right player (paddle 1) up/down, [6]/[7] left player (paddle 0) up/down, SELECT / RESET of either player
    BUTTON_MAP: Tuple[int, ...] = (2, 0, 13, 14, 7, 8, 5, 6, 0, 0, 13, 14, 0, 0, 0, 1)

    @property
    def GAME_PLAYABLE_HEIGHT(self) -> int:
        return self.GAME_BOTTOM - self.GAME_TOP

    @property
    def TOURNAMENT_SIZE(self) -> int:          # config.py:49
        return self.POPULATION_SIZE // 4

    @property
    def HALL_OF_FAME_AMOUNT(self) -> int:      # config.py:50
        return self.TOURNAMENT_SIZE

    def gene_size(self) -> int:                # utils.calculate_gene_size, utils.py:128-136
        b = 1 if self.BIAS else 0
        return sum((self.NETWORK_SHAPE[i] + b) * self.NETWORK_SHAPE[i + 1] for i in range(len(self.NETWORK_SHAPE) - 1))

    def blank_action(self):                    # BLANK_ACTION, config.py:21-23
        a = [0] * 16
        a[self.LEFT_PLAYER_START_BUTTON] = 1
        a[self.RIGHT_PLAYER_START_BUTTON] = 1
        return a

    def replace(self, **kw) -> "Config":
        return replace(self, **kw)

    def to_c(self) -> _lib.NgpConfig:
        assert (self.GAME_TOP, self.GAME_BOTTOM, self.GAME_WIDTH) == (34, 194, 160), "frame geometry is fixed by the cartridge"
        c = _lib.NgpConfig()
        c.n_layers = len(self.NETWORK_SHAPE)
        for i, n in enumerate(self.NETWORK_SHAPE):
            c.nodes[i] = n
        c.bias = 1 if self.BIAS else 0
        c.games_to_play = self.GAMES_TO_PLAY
        c.win_score = self.WIN_SCORE
        c.timeout_thresh = self.TIMEOUT_THRESH
        c.schedule = self.SCHEDULE
        c.max_frames = self.MAX_FRAMES
        c.time_scaler = self.TIME_SCALER
        c.scaled_paddle_height = self.SCALED_PADDLE_HEIGHT
        for i in range(3):
            c.ball_colour[i] = self.BALL_COLOUR[i]
            c.left_colour[i] = self.LEFT_GUY_COLOUR[i]
            c.right_colour[i] = self.RIGHT_GUY_COLOUR[i]
        c.cxpb = self.CROSSOVER_BLEND_PROBABILITY
        c.cx_alpha = self.CROSSOVER_BLEND_ALPHA
        c.mutpb = self.GAUSSIAN_MUTATION_PROBABILITY
        c.mut_mu = self.GAUSSIAN_MUTATION_MEAN
        c.mut_sigma = self.GAUSSIAN_MUTATION_SIGMA
        c.mut_indpb = self.PROBABILITY_OF_MUTATING_A_SINGLE_GENE
        c.tournament_size = self.TOURNAMENT_SIZE
        c.core = self.CORE
        assert len(self.BUTTON_MAP) == 16
        for i, code in enumerate(self.BUTTON_MAP):
            c.button_map[i] = code
        return c
