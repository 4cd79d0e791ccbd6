from .safety_gym import (SafetyGymStateScorer, ScorerEnvironment, make_sensor_offset_table,
                         POINTGOAL1_SENSORS, POINTSIMPLEGOAL1_SENSORS, DEFAULT_CONFIG)
