"""The ``__main__`` body of the reference's run_imitator.py (lines 215-241), verbatim, as a string: executed by
tests/test_run_imitator_gpu.py against this repo's ``Imitator`` and compared with the reference file by
tests/test_run_imitator_cpu.py wherever /root/reference exists.  ``post_tune`` stays False (fine-tuning needs the backward
pass), so ``adaptive_personalize`` is never entered."""

BODY = '''    # meta imitator
    test_opt = TestOptions().parse()

    if test_opt.ip:
        visualizer = VisdomVisualizer(env=test_opt.name, ip=test_opt.ip, port=test_opt.port)
    else:
        visualizer = None

    # set imitator
    imitator = Imitator(test_opt)

    if test_opt.post_tune:
        adaptive_personalize(test_opt, imitator, visualizer)

    imitator.personalize(test_opt.src_path, visualizer=visualizer)
    print('\\n\\t\\t\\tPersonalization: completed...')

    if test_opt.save_res:
        pred_output_dir = mkdir(os.path.join(test_opt.output_dir, 'imitators'))
        pred_output_dir = clear_dir(pred_output_dir)
    else:
        pred_output_dir = None

    print('\\n\\t\\t\\tImitating `{}`'.format(test_opt.tgt_path))
    tgt_paths = scan_tgt_paths(test_opt.tgt_path, itv=1)
    imitator.inference(tgt_paths, tgt_smpls=None, cam_strategy='smooth',
                       output_dir=pred_output_dir, visualizer=visualizer, verbose=True)
'''
