"""Names existing agents import from the reference (luxai2021/game/constants.py:4-96,
luxai2021/game/game_constants.json:18-57).  They are the interface, not engine inputs: the CUDA engine
compiles the default parameter set in (csrc/lux_step_core.cuh), so `configs["parameters"]` is carried for
agents that read it and `Game` refuses a parameter set that differs from it.
"""


class Constants:
    class AGENT_TYPE:
        AGENT = "agent"
        LEARNING = "learning"

    class INPUT_CONSTANTS:
        RESEARCH_POINTS, RESOURCES, UNITS, CITY, CITY_TILES, ROADS, DONE = "rp", "r", "u", "c", "ct", "ccd", "D_DONE"

    class DIRECTIONS:
        NORTH, WEST, SOUTH, EAST, CENTER = "n", "w", "s", "e", "c"

    class UNIT_TYPES:
        WORKER, CART = 0, 1

    class TEAM:
        A, B = 0, 1

    class RESOURCE_TYPES:
        WOOD, URANIUM, COAL = "wood", "uranium", "coal"

    class MAP_TYPES:
        EMPTY, RANDOM, DEBUG = "empty", "random", "debug"

    class ACTIONS:
        MOVE, RESEARCH, BUILD_WORKER, BUILD_CART, BUILD_CITY, TRANSFER, PILLAGE = "m", "r", "bw", "bc", "bcity", "t", "p"


GAME_CONSTANTS = {
    "UNIT_TYPES": {"WORKER": 0, "CART": 1},
    "RESOURCE_TYPES": {"WOOD": "wood", "COAL": "coal", "URANIUM": "uranium"},
    "DIRECTIONS": {"NORTH": "n", "WEST": "w", "EAST": "e", "SOUTH": "s", "CENTER": "c"},
    "PARAMETERS": {
        "DAY_LENGTH": 30, "NIGHT_LENGTH": 10, "MAX_DAYS": 360,
        "LIGHT_UPKEEP": {"CITY": 23, "WORKER": 4, "CART": 10},
        "WOOD_GROWTH_RATE": 1.025, "MAX_WOOD_AMOUNT": 500,
        "CITY_BUILD_COST": 100, "CITY_ADJACENCY_BONUS": 5,
        "RESOURCE_CAPACITY": {"WORKER": 100, "CART": 2000},
        "WORKER_COLLECTION_RATE": {"WOOD": 20, "COAL": 5, "URANIUM": 2},
        "RESOURCE_TO_FUEL_RATE": {"WOOD": 1, "COAL": 10, "URANIUM": 40},
        "RESEARCH_REQUIREMENTS": {"COAL": 50, "URANIUM": 200},
        "CITY_ACTION_COOLDOWN": 10, "UNIT_ACTION_COOLDOWN": {"CART": 3, "WORKER": 2},
        "MAX_ROAD": 6, "MIN_ROAD": 0, "CART_ROAD_DEVELOPMENT_RATE": 0.75, "PILLAGE_RATE": 0.5,
    },
}

LuxMatchConfigs_Default = {
    "mapType": Constants.MAP_TYPES.RANDOM,
    "storeReplay": True,
    "seed": None,
    "debug": False,
    "debugDelay": 500,
    "runProfiler": False,
    "compressReplay": False,
    "debugAnnotations": False,
    "statefulReplay": False,
    "parameters": GAME_CONSTANTS["PARAMETERS"],
}
