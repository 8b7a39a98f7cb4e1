"""``Team`` and ``Message`` of the reference's intra-team communication API (mate/utils.py:274-307), for callers that
do not have the reference package; ``MultiAgentTracking.send_messages`` accepts either these or the reference's."""

import enum
from dataclasses import dataclass
from typing import Any, Optional

__all__ = ['Team', 'Message']


class Team(enum.Enum):
    """mate/utils.py:274-278."""

    CAMERA = 0
    TARGET = 1


@dataclass
class Message:
    """One message between agents of the same team (mate/utils.py:281-307)."""

    sender: int
    recipient: Optional[int]      # None = broadcast to all teammates
    content: Any
    team: Team
    broadcasting: bool = False

    def __contains__(self, name):
        return name in self.content

    def __getitem__(self, name):
        return self.content[name]

    def __setitem__(self, name, value):
        self.content[name] = value
