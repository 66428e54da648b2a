package PhyloProf::B200;

# Perl host of the B200-native PPP scan: loads the XS binding of libppp_b200.so (include/ppp_b200.h), packs
# PhyloProf::Profile objects into bit planes and scores them on the GPU.  There is no CPU fallback: without the
# shared objects or without an sm_100 device every entry point croaks.
#
#   my $gpu  = PhyloProf::B200->new(-device => 0);                 # after any fork (bin/ppp.pl:910 forks workers)
#   my $gpu8 = PhyloProf::B200->new(-devices => "0-7");            # one process, eight GPUs
#   my $hits = $gpu->score_bits(\@universe, \@queries, \@targets, [-filter => 'topk', -k => 100]);
#
# Data model (SURVEY.md section 8d): @universe is the ordered taxon list (ascending taxon id as data/taxdis.txt and
# bin/hmm3profile.pl:175-181 print it); a query is a PhyloProf::Profile (taxon => 0/1 hash, Profile.pm:433-514), a
# target is a set of taxa whose list order is the universe order.
use strict;
use warnings;
use Carp;
require DynaLoader;
our @ISA     = qw(DynaLoader);
our $VERSION = '0.1';

sub dl_load_flags { 0x01 }    # RTLD_GLOBAL not needed; keep lazy binding

my $loaded = 0;

sub _load {
    return if $loaded;
    ( my $dir = __FILE__ ) =~ s{lib/PhyloProf/B200\.pm$}{};
    my $so = $ENV{PPP_B200_XS} || "${dir}blib/PPPB200.so";
    croak "PhyloProf::B200: $so not built (make -C perl); there is no pure-Perl or CPU path" unless -e $so;
    my $lib = DynaLoader::dl_load_file( $so, 0 ) or croak "PhyloProf::B200: " . DynaLoader::dl_error();
    my $sym = DynaLoader::dl_find_symbol( $lib, "boot_PhyloProf__B200" ) or croak DynaLoader::dl_error();
    my $xs  = DynaLoader::dl_install_xsub( "PhyloProf::B200::_bootstrap", $sym, $so );
    &$xs("PhyloProf::B200");
    $loaded = 1;
}

sub new {
    my ( $class, %a ) = @_;
    _load();
    # -device => N (one GPU) or -devices => [N, M, ...] / "0-3" / "0,2" (one process drives them all: the library shards the
    # targets over the devices, scores them concurrently and assembles the results; replaces bin/ppp.pl:895-922's forked clients)
    my $devs = $a{-devices};
    $devs = parse_devices($devs) if defined $devs && !ref $devs;
    $devs ||= [ $a{-device} || 0 ];
    my $self = bless { h => _init($devs), ndev => scalar(@$devs) }, $class;
    return $self;
}

sub parse_devices {    # "0-7", "0,2,5", "3" -> [ordinals]
    my $spec = shift;
    my @d;
    foreach my $part ( split /,/, $spec ) {
        if    ( $part =~ /^(\d+)-(\d+)$/ ) { push @d, ( $1 .. $2 ) }
        elsif ( $part =~ /^\d+$/ )         { push @d, $part + 0 }
        else                               { croak "PhyloProf::B200: bad device list '$spec'" }
    }
    croak "PhyloProf::B200: empty device list" unless @d;
    return \@d;
}

sub DESTROY { my $s = shift; _destroy( $s->{h} ) if $s->{h}; $s->{h} = 0 }

# taxon list -> index map
sub _index { my $u = shift; my %i; @i{@$u} = ( 0 .. $#$u ); return \%i }

sub _plane {
    my ( $wpr, $idx, $taxa ) = @_;
    my $v = "\0" x ( $wpr * 4 );
    foreach my $t (@$taxa) {
        my $g = $idx->{$t};
        next unless defined $g;
        vec( $v, $g, 1 ) = 1;    # vec bit g of the string = bit (g & 7) of byte g >> 3 = bit (g & 31) of LE word g >> 5
    }
    return $v;
}

# score every query against every target; targets are ORDERED hit lists in their own order, duplicates and foreign taxa
# allowed (BlastResult.pm:192-237 / value-sorted HMMERResult hash): de-duplicated here, raw positions kept (ppp.pm:169-231)
#
# -dup => 1 is the reference's -dup flag (ppp.pm:189-200, 214-217).  Its `$seen{key}` is ONE literal-key counter for the
# whole list: the first repeated entry of a list is counted again like a new taxon (kept here as a second entry), and
# from the second repeated entry on `yes` is frozen while `number` only grows, so no later step can improve and the list
# ends there.  The device drops the candidate that would exceed |Y| (`last if $yes > $max_yes`, ppp.pm:222-224).
sub score_ranked {
    my ( $self, $universe, $queries, $lists, %o ) = @_;
    my $idx = _index($universe);
    my ( @gi, @raw, @off ) = ();
    push @off, 0;
    foreach my $l (@$lists) {
        my ( %seen, $repeats );
        croak "target list of " . scalar(@$l) . " positions exceeds 65535" if @$l > 65535;    # raw positions are uint16
        for my $i ( 0 .. $#$l ) {
            if ( $seen{ $l->[$i] }++ ) {
                next unless $o{-dup};
                last if ++$repeats >= 2;
            }
            push @gi, ( exists $idx->{ $l->[$i] } ? $idx->{ $l->[$i] } : 0xFFFF );
            push @raw, $i + 1;
        }
        push @off, scalar @gi;
    }
    _set_targets_ranked( $self->{h}, pack( "S*", @gi ), pack( "S*", @raw ), pack( "Q*", @off ), scalar(@$lists), scalar(@$universe) );
    return $self->_score_dup( $o{-dup}, $universe, $idx, $queries, %o );
}

sub _score_dup {
    my ( $self, $dup, @rest ) = @_;
    _set_option( $self->{h}, "dup", $dup ? 1 : 0 );
    my $hits = eval { $self->_score_queries(@rest) };
    my $err  = $@;
    _set_option( $self->{h}, "dup", 0 );
    die $err if $err;
    return $hits;
}

# score every query against the hit lists cached in `.bla` files (one target per file; bin/ppp.pl:606-673 format): the
# files are parsed, de-duplicated and mapped to the universe natively (ppp_read_bla_flags), no Perl-side parsing;
# -dup => 1 as in score_ranked
sub score_bla_files {
    my ( $self, $universe, $queries, $paths, %o ) = @_;
    my $idx = _index($universe);
    _set_targets_bla( $self->{h}, $paths, $universe, $o{-threads} || 0, $o{-dup} ? 1 : 0 );    # 1 = PPP_BLA_DUP
    return $self->_score_dup( $o{-dup}, $universe, $idx, $queries, %o );
}

# score every query (PhyloProf::Profile or {taxon=>0/1}) against every target ([taxa] in universe order semantics)
sub score_bits {
    my ( $self, $universe, $queries, $targets, %o ) = @_;
    my $G = scalar @$universe;
    croak "universe of $G genomes outside [1, 8192]" if $G < 1 || $G > 8192;
    my $idx = _index($universe);
    my $wpr = words_per_row($G);
    my $tb  = join '', map { _plane( $wpr, $idx, $_ ) } @$targets;
    _set_targets_bits( $self->{h}, $tb, scalar(@$targets), $G );
    return $self->_score_queries( $universe, $idx, $queries, %o );
}

# after a -filter => 'thresh' call: per query [first row, rows, marker row | -1, CUTOFF in thousandths | undef, status] -- the
# ppp.pl -m <slope> marker and ppp_cutoff.pl -s <slope> CUTOFF of each query's (already sorted) rows, computed on the device
# (bin/ppp.pl:473-506, bin/ppp_cutoff.pl:130-148).  status != 0: format that query's rows in Perl (include/ppp_b200.h).
sub printed_cutoffs {
    my ( $self, $n_queries, $slope ) = @_;    # $n_queries is checked against the run's own query count by the XS layer
    return [ _printed_cutoffs( $self->{h}, $n_queries, $slope ) ];
}

sub _score_queries {
    my ( $self, $universe, $idx, $queries, %o ) = @_;
    my $G   = scalar @$universe;
    my $wpr = words_per_row($G);
    my ( $pb, $yb, @p ) = ( '', '' );
    foreach my $q (@$queries) {
        my $h = ref($q) eq 'HASH' ? $q : $q->get_hash();
        my @present = grep { exists $idx->{$_} } keys %$h;
        my @yes     = grep { $h->{$_} != 0 } @present;
        $pb .= _plane( $wpr, $idx, \@present );
        $yb .= _plane( $wpr, $idx, \@yes );
        # p = get_yes / get_total of the WHOLE profile, taxa outside the universe included (ppp.pm:138-142); a false -prob
        # counts as absent (ppp.pm:100-102)
        my ( $yes_all, $total ) = ref($q) eq 'HASH' ? ( scalar( grep { $_ != 0 } values %$h ), scalar( keys %$h ) ) : ( $q->get_yes(), $q->get_total() );
        croak "Illegal division by zero" unless $o{-prob} || $o{-probs} || $total;    # ppp.pm:141
        push @p, $o{-prob} ? $o{-prob} : ( $total ? $yes_all / $total : 0 );
    }
    @p = @{ $o{-probs} } if $o{-probs};
    my %mode = ( all => 0, thresh => 1, topk => 2 );
    my $m = $mode{ $o{-filter} || 'all' };
    croak "unknown -filter" unless defined $m;
    my @hits = _score( $self->{h}, $pb, $yb, pack( "d*", @p ), scalar(@$queries), $m, $o{-k} || 0, $o{-thresh} || 0 );
    return \@hits;    # [q, t, score, yes, number, depth]
}

1;
