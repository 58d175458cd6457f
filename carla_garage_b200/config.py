"""Hyper-parameter surface of the TransFuser++ hot path: the subset of the reference's ``GlobalConfig``
(team_code/config.py:10-598) that LidarCenterNet / TransfuserBackbone / the LiDAR voxeliser read, same names, same
defaults (line numbers cited).  The product modules accept EITHER this object or the reference's own GlobalConfig
(any attribute bag works) — this class exists so tests / bench run where /root/reference and ``carla`` do not."""


class GlobalConfig:

  def __init__(self):
    # autopilot speeds (config.py:33-35,148)
    self.target_speed_slow, self.target_speed_fast, self.target_speed_walker = 5.0, 8.0, 2.0
    self.target_speeds = [0.0, self.target_speed_walker, self.target_speed_slow, self.target_speed_fast]
    # sensors (config.py:94-107)
    self.camera_pos = [-1.5, 0.0, 2.0]
    self.camera_rot_0 = [0.0, 0.0, 0.0]
    self.camera_width, self.camera_height, self.camera_fov = 1024, 256, 110
    # dataloader (config.py:111-131)
    self.carla_fps = 20
    self.data_save_freq = 5
    self.seq_len = self.img_seq_len = self.lidar_seq_len = 1
    self.lidar_resolution_width = self.lidar_resolution_height = 256
    self.pixels_per_meter = 4.0
    self.hist_max_per_pixel = 5
    self.lidar_split_height = 0.2
    self.use_ground_plane = False
    self.min_x, self.max_x, self.min_y, self.max_y = -32, 32, -32, 32
    self.min_z_projection, self.max_z_projection = -10, 14
    self.max_height_lidar = 100.0  # config.py:481
    # class weights (config.py:158-164)
    self.target_speed_weights = [0.866605263873406, 7.4527377240841775, 1.2281629310898465, 0.5269622904065803]
    self.semantic_weights = [1.0] * 7
    self.bev_semantic_weights = [1.0] * 11
    # training (config.py:169-259)
    self.lr = 0.0003
    self.batch_size = 32
    self.epochs = 31
    self.sync_batch_norm = False
    self.zero_redundancy_optimizer = 1
    self.detect_boxes = 1
    self.backbone = 'transFuser'
    self.use_velocity = 1
    self.image_architecture = 'regnety_032'
    self.lidar_architecture = 'regnety_032'
    self.use_controller_input_prediction = True
    self.label_smoothing_alpha = 0.1
    self.use_focal_loss = False
    self.use_amp = 0
    self.use_grad_clip = 0
    self.use_bev_semantic = True
    self.use_depth = True
    self.detailed_loss_weights = {
        'loss_wp': 1.0, 'loss_target_speed': 1.0, 'loss_checkpoint': 1.0, 'loss_semantic': 1.0,
        'loss_bev_semantic': 1.0, 'loss_depth': 1.0, 'loss_center_heatmap': 1.0, 'loss_wh': 1.0, 'loss_offset': 1.0,
        'loss_yaw_class': 1.0, 'loss_yaw_res': 1.0, 'loss_velocity': 1.0, 'loss_brake': 1.0, 'loss_forcast': 0.2,
        'loss_selection': 0.0,
    }
    self.use_speed_weights = True  # config.py:262
    self.use_label_smoothing = False  # config.py:266
    # controller (config.py:250-287)
    self.brake_speed, self.brake_ratio, self.clip_delta, self.clip_throttle = 0.4, 1.1, 0.25, 0.75
    self.aim_distance_fast, self.aim_distance_slow, self.aim_distance_threshold = 3.0, 2.25, 5.5
    self.turn_kp, self.turn_ki, self.turn_kd, self.turn_n = 1.25, 0.75, 0.3, 20
    self.speed_kp, self.speed_ki, self.speed_kd, self.speed_n = 5.0, 0.5, 1.0, 20
    self.debug = False
    # detector (config.py:307-322)
    self.bb_confidence_threshold = 0.3
    self.iou_treshold_nms = 0.2  # config.py:492 (the reference's spelling)
    self.num_dir_bins = 12
    self.top_k_center_keypoints = 100
    self.center_net_max_pooling_kernel = 3
    self.bb_input_channel = 64
    self.num_bb_classes = 4
    # model (config.py:327-366)
    self.gru_hidden_size = 64
    self.gru_input_size = 256
    self.img_vert_anchors = self.camera_height // 32
    self.img_horz_anchors = self.camera_width // 32
    self.lidar_vert_anchors = self.lidar_resolution_height // 32
    self.lidar_horz_anchors = self.lidar_resolution_width // 32
    self.perspective_downsample_factor = 1
    self.bev_features_chanels = 64
    self.bev_down_sample_factor = 4
    self.bev_upsample_factor = 2
    self.block_exp, self.n_layer, self.n_head = 4, 2, 4
    self.embd_pdrop = self.resid_pdrop = self.attn_pdrop = 0.1
    self.gpt_linear_layer_init_mean, self.gpt_linear_layer_init_std = 0.0, 0.02
    self.gpt_layer_norm_init_weight = 1.0
    self.predict_checkpoint_len = 10
    self.normalize_imagenet = True
    self.use_wp_gru = False
    self.use_semantic = True
    self.num_semantic_classes = 7
    self.num_bev_semantic_classes = 11  # len(bev_converter), config.py:420-449
    self.deconv_channel_num_0, self.deconv_channel_num_1, self.deconv_channel_num_2 = 128, 64, 32  # config.py:451-453
    self.deconv_scale_factor_0, self.deconv_scale_factor_1 = 4, 8  # config.py:456-458
    self.use_discrete_command = True
    self.add_features = True
    self.transformer_decoder_join = True
    self.num_transformer_decoder_layers = 6
    self.num_decoder_heads = 8
    self.bev_grid_height_downsample_factor = 1.0
    self.image_u_net_output_features = 512  # config.py:463-464 (bev_encoder backbone)
    self.bev_latent_dim = 32
    self.wp_dilation = 1
    self.extra_sensor_channels = 128
    self.use_tp = True
    self.learn_origin = 1  # config.py:192
    self.tp_attention = False
    self.multi_wp_output = False
    self.pred_len = int(2.0 * self.carla_fps) // self.data_save_freq
    self.use_plant = False

  def initialize(self, **kwargs):
    """config.py:546-548: every keyword becomes an attribute."""
    for k, v in kwargs.items():
      setattr(self, k, v)
