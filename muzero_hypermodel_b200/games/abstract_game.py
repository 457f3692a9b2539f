"""The AbstractGame contract (games/abstract_game.py:4-105): step / legal_actions / to_play / reset."""
from abc import ABC, abstractmethod


class AbstractGame(ABC):
    @abstractmethod
    def __init__(self, seed=None):
        pass

    @abstractmethod
    def step(self, action):
        """Apply `action`; returns (observation, reward, done)."""

    def to_play(self):
        """Current player, an element of config.players."""
        return 0

    @abstractmethod
    def legal_actions(self):
        """List of legal actions (subset of config.action_space)."""

    @abstractmethod
    def reset(self):
        """Start a new game; returns the initial observation."""

    def close(self):
        pass

    def render(self):
        raise NotImplementedError

    def human_to_action(self):
        choice = input(f"Enter the action to play for the player {self.to_play()}: ")
        while int(choice) not in self.legal_actions():
            choice = input("Illegal action. Enter another action : ")
        return int(choice)

    def expert_agent(self):
        raise NotImplementedError

    def action_to_string(self, action_number):
        return str(action_number)
