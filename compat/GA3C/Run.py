from rl_collision_avoidance_b200.ga3c.Run import main

if __name__ == '__main__':
    main()
