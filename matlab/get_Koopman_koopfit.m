function [ koopData , K ] = get_Koopman_koopfit( obj , snapshotPairs , varargin )
%get_Koopman_koopfit: drop-in body for Ksysid.get_Koopman (Ksysid.m:987-1092) that runs the
% lift loop, Px'Px / Px'Py, `\` and the L1-ball QP on a B200 through koopfit_mex (libkoopfit.so).
% Same inputs, same koopData fields.  UNTESTED IN MATLAB (no MATLAB/Octave in the build image);
% the C ABI it calls is the tested artefact (see INTEGRATION.md).
    desc.types   = obj.obs_type;                 % cellstr as given to the constructor (Ksysid.m:20)
    desc.degrees = obj.obs_degree;               % Ksysid.m:21
    if isfield( obj.params , 'gauss_centres' )   % zeta0 of EVERY def_gaussianLift call, appended after Ksysid.m:803:
        desc.centres = obj.params.gauss_centres; %   obj.params.gauss_centres = [ obj.params.gauss_centres , zeta0 ];
    else
        desc.centres = [];
    end
    if ~isempty( obj.dim_red ) && ~obj.dim_red   % identity econ lifts (Ksysid.m:1437-1494)
        desc.pcs = [];
    else
        desc.pcs = obj.basis.pcs;                % Ksysid.m:1507-1510
    end
    if length(varargin) == 1 && ~isempty( varargin{1} )
        lasso = varargin{1};
    else
        lasso = 1e4;                             % Ksysid.m:994,999
    end
    opts.least_squares    = all( obj.lasso >= 1e6 );   % branch on the PROPERTY (Ksysid.m:1068)
    opts.t                = lasso * obj.params.N;      % Ksysid.m:996
    opts.delay_constraint = strcmp( obj.model_type , 'linear' ) && obj.params.nd >= 1;   % Ksysid.m:1139-1164
    opts.n  = obj.params.n;
    opts.nd = obj.params.nd;
    want_reg = strcmp( obj.model_type , 'linear' );    % get_model needs koopData.Px / Py (Ksysid.m:1206-1216)

    disp('Finding Koopman operator approximation...');              % Ksysid.m:1002
    if obj.loaded && isfield( snapshotPairs , 'w' )                  % Ksysid.m:1006-1011
        w = snapshotPairs.w;   nw = obj.params.nw;
    else
        w = [];                nw = 0;
    end
    % one call, every visible GPU (kf_fit_multi): shards + one NCCL all-reduce inside libkoopfit.so
    [ Kall , info , Px , Py ] = koopfit_mex( 'fit' , snapshotPairs.alpha , snapshotPairs.beta , snapshotPairs.u , ...
                                             obj.model_type , desc , opts , want_reg , w );
    K = Kall(:,:,1);
    koopData.K = K;                                                  % Ksysid.m:1084
    if want_reg
        N = obj.params.N;
        koopData.Px = Px( : , 1:N*(nw+1) );                          % Ksysid.m:1085-1086
        koopData.Py = Py( : , 1:N*(nw+1) );
    end
    koopData.u = snapshotPairs.u;                                    % Ksysid.m:1087
    if nw > 0
        koopData.w = w;                                              % Ksysid.m:1088-1090
    end
    koopData.alpha = snapshotPairs.alpha;                            % Ksysid.m:1091
    koopData.info = info;                                            % rank, method, timings (new)
end
